#!/usr/bin/env python
"""bench.py -- kanzi block pipeline on B200: encode+decode MB/s, bit-exact stream.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    -t BWT+RANK+ZRLT -e ANS0 -b 4m, 1 GiB synth_compressible(seed 2), block-sharded
    round-robin over N GPUs (strong scaling: the 1 GiB is fixed).
A "step" = one encode pass + one decode pass over the whole input.

  value   device-resident: blocks already in HBM when the timed region starts;
          encode (all ranks) -> NCCL gather of block payloads to rank 0 -> bit
          assembly of the kanzi stream on rank 0 -> decode (all ranks).
  e2e     same metric through the public API with HOST buffers: pinned-host -> device
          copies of the input, device -> host of the compressed stream and of the
          decoded bytes inside the timed region (knz_compress / knz_decompress at N=1).
  roofline  dominant kernel of the step (inverse RANK, one launch) against the HBM roofline
          with SURVEY.md §8(d) algorithmic bytes (2n per block) and the DRAM traffic of the
          committed ncu capture; every stage, incl. BWT forward (2n+25) and the rANS kernels
          alone (m + e bytes), is under roofline_stages.
  cpu_baseline / --impl reference: the unmodified reference (oracle/_ref, built from
          /root/reference by oracle/Makefile) with all host threads on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "kanzi-cpp_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

TRANSFORM, ENTROPY, BLOCK = "BWT+RANK+ZRLT", "ANS0", 4 << 20
METRIC = "encode+decode MB/s (BWT+RANK+ZRLT / ANS0, -b 4m, 1 GiB synthetic), bit-exact stream"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def cpu_reference_run(data, jobs, reps=1):
    """Times the unmodified reference (oracle/_ref) encode+decode on `data`."""
    import hashlib
    from oracle.oracle import Ref
    ref = Ref.load()
    if ref is None:
        return None
    best_e = best_d = 1e30
    comp = None
    for _ in range(reps):
        t = time.perf_counter()
        comp = ref.stream_compress(data, TRANSFORM, ENTROPY, BLOCK, jobs=jobs)
        best_e = min(best_e, time.perf_counter() - t)
        t = time.perf_counter()
        back, rc = ref.stream_decompress(comp, data.size, jobs=jobs)
        best_d = min(best_d, time.perf_counter() - t)
        assert rc == 0 and back.size == data.size
    return {"enc_s": best_e, "dec_s": best_d, "comp": int(comp.size),
            "sha256": hashlib.sha256(comp.tobytes()).hexdigest()}


def run_reference(args, rank, world):
    """Reference arm: the unmodified reference's own multi-threaded CPU path (oracle/_ref), the FULL
    workload every step (same config as the GPU arm), in-memory streams, all host threads."""
    if rank != 0:
        return
    import synth
    cores = os.cpu_count() or 1
    jobs = max(1, min(63, cores))  # 64 trips a reference bug (jobsPerTask[63] = 0 once >= 63 blocks are queued)
    size = (args.size // BLOCK) * BLOCK
    data = synth.synth_compressible(size, 2)
    from oracle.oracle import Ref
    if Ref.load() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libkanzi_ref.so not built"}))
        return
    if args.warmup > 0:
        cpu_reference_run(data[: 32 << 20], jobs)
    # the whole --steps/--warmup run must end within a few minutes: cap the timed steps by a time budget
    tot_e = tot_d = 0.0
    done = 0
    sha = None
    t_start = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_reference_run(data, jobs)
        tot_e += r["enc_s"]
        tot_d += r["dec_s"]
        sha = r["sha256"]
        comp_bytes = r["comp"]
        done += 1
        if time.perf_counter() - t_start > 150:
            break
    per_step = (tot_e + tot_d) / done
    value = size / per_step / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "encode_MBps": size * done / tot_e / 1e6, "decode_MBps": size * done / tot_d / 1e6,
        "config": {"workload": f"-t {TRANSFORM} -e {ENTROPY} -b 4m, {size >> 20} MiB synth_compressible(seed 2)",
                   "blocks": size // BLOCK, "jobs": jobs,
                   "build": "unmodified reference sources, g++ -O3 -march=x86-64-v3 (the library travels between "
                            "hosts, so not -march=native)"},
        "stream_sha256": sha, "compressed_bytes": comp_bytes,
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": jobs, "kind": "reference",
                         "sample": f"the full {size >> 20} MiB workload per step, jobs={jobs}, in-memory streams"},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import synth
    from kanzi_b200 import Context, E_IDS, _ptr, sharded

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    size = (args.size // BLOCK) * BLOCK
    nblocks = size // BLOCK
    assert nblocks % world == 0, "block count must divide over the ranks"
    data = synth.synth_compressible(size, 2)
    my = list(range(rank, nblocks, world))  # round-robin sharding (north_star)
    nb = len(my)
    host_in = torch.from_numpy(data).view(nblocks, BLOCK)[my].contiguous().pin_memory()
    batch = min(args.batch, nb)
    ctx = Context(local_rank, BLOCK, batch)
    L = ctx.lib
    ostride = (BLOCK + BLOCK // 4 + 4096 + 255) // 256 * 256
    d_in = torch.empty((nb, BLOCK), dtype=torch.uint8, device=dev)
    d_in.copy_(host_in)
    d_blk = torch.zeros((nb, ostride), dtype=torch.uint8, device=dev)
    d_bits = torch.zeros(nb, dtype=torch.int64, device=dev)
    d_dec = torch.empty((nb, BLOCK), dtype=torch.uint8, device=dev)
    lens = np.full(nb, BLOCK, dtype=np.int32)
    ttype = ctx.transform_type(TRANSFORM)
    etype = E_IDS[ENTROPY]
    hdr = np.zeros(32, dtype=np.uint8)
    hdr_bytes = L.knz_stream_header(ttype, etype, BLOCK, size, _ptr(hdr))
    stream_cap = size + size // 4 + 65536
    d_stream = torch.zeros(stream_cap if rank == 0 else 16, dtype=torch.uint8, device=dev)
    stage_ms = {"enc": None, "dec": None}
    info = {}

    def encode_dev():
        sharded.encode_shard(ctx, ttype, etype, BLOCK, d_in, lens, BLOCK, d_blk, d_bits)
        stage_ms["enc"] = ctx.timings()

    def gather_and_assemble():
        """Block payloads -> rank 0 over NCCL, then the bit-concatenation kernel lays them
        down at their bit offsets in stream order (kanzi_b200/sharded.py)."""
        res = sharded.gather_blocks(d_blk, d_bits, rank, world)
        if res is None:
            return None
        blk, bits_t = res
        d_stream.zero_()
        torch.cuda.synchronize()
        end = sharded.assemble_stream(ctx, blk, bits_t, d_stream, 8 * hdr_bytes)
        return (end + 8 + 7) // 8

    def decode_dev():
        out_lens = sharded.decode_shard(ctx, ttype, etype, BLOCK, d_blk, d_bits.cpu().numpy().astype(np.uint64), d_dec)
        assert (out_lens == BLOCK).all()
        stage_ms["dec"] = ctx.timings()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        encode_dev()
        nbytes = gather_and_assemble()
        t_mid = time.perf_counter()
        decode_dev()
        return nbytes, t_mid

    # ---- warm-up + correctness of what is being timed
    comp_bytes = None
    for _ in range(max(1, args.warmup)):
        comp_bytes, _ = step_dev()
    barrier()
    assert bool((d_dec == d_in).all().item()), "decode(encode(x)) != x"
    if rank == 0:
        import hashlib
        stream = d_stream[:comp_bytes].cpu().numpy()
        stream[:hdr_bytes] = hdr[:hdr_bytes]
        info["stream_sha256"] = hashlib.sha256(stream.tobytes()).hexdigest()
        info["compressed_bytes"] = int(comp_bytes)

    # ---- timed: device-resident
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launches
    barrier()
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    lib_stream = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    ev0.record(lib_stream)
    t0 = time.perf_counter()
    enc_wall = dec_wall = 0.0
    for _ in range(args.steps):
        ts = time.perf_counter()
        _, t_mid = step_dev()
        torch.cuda.synchronize()
        te = time.perf_counter()
        enc_wall += t_mid - ts
        dec_wall += te - t_mid
    barrier()
    ev1.record(lib_stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    t_local = torch.tensor([max(dev_ms / 1e3, wall), enc_wall, dec_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    total_s, enc_s, dec_s = [float(x) for x in t_local.tolist()]
    per_step = total_s / args.steps

    # ---- e2e through the public API with host buffers
    e2e = None
    if world == 1:
        host_full = torch.from_numpy(data).pin_memory().numpy()
        out_comp = torch.empty(stream_cap, dtype=torch.uint8).pin_memory().numpy()
        out_plain = torch.empty(size, dtype=torch.uint8).pin_memory().numpy()
        comp = ctx.compress(host_full, TRANSFORM, ENTROPY, BLOCK, out=out_comp)  # warm
        torch.cuda.synchronize()
        t = time.perf_counter()
        e_enc = e_dec = 0.0
        for _ in range(args.steps):
            ts = time.perf_counter()
            comp = ctx.compress(host_full, TRANSFORM, ENTROPY, BLOCK, out=out_comp)
            tm = time.perf_counter()
            back = ctx.decompress(comp, size, out=out_plain)
            te = time.perf_counter()
            e_enc += tm - ts
            e_dec += te - tm
        torch.cuda.synchronize()
        e_s = (time.perf_counter() - t) / args.steps
        assert back.size == size and np.array_equal(back[: 1 << 20], data[: 1 << 20])
        e2e = {"value": size / e_s / 1e6, "unit": "MB/s",
               "h2d_bytes_per_step": int(size + comp.size), "d2h_bytes_per_step": int(comp.size + size),
               "encode_MBps": size * args.steps / e_enc / 1e6, "decode_MBps": size * args.steps / e_dec / 1e6,
               "api": "knz_compress + knz_decompress (host pinned buffers)"}
    else:
        # host -> device of the rank's blocks, device path, stream and decoded blocks back to the host
        host_out = torch.empty((nb, BLOCK), dtype=torch.uint8).pin_memory()
        host_stream = torch.empty(stream_cap if rank == 0 else 16, dtype=torch.uint8).pin_memory()
        barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            d_in.copy_(host_in, non_blocking=True)
            nbytes, _ = step_dev()
            if rank == 0:
                host_stream[:nbytes].copy_(d_stream[:nbytes], non_blocking=True)
            host_out.copy_(d_dec, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e_s = torch.tensor([(time.perf_counter() - t) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": size / float(e_s.item()) / 1e6, "unit": "MB/s", "h2d_bytes_per_step": int(size),
               "d2h_bytes_per_step": int(size + (comp_bytes or 0)),
               "api": "knz_encode_blocks_dev/knz_decode_blocks_dev per rank + NCCL gather + knz_assemble_stream_dev"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (stage level, SURVEY.md §8(d) algorithmic bytes), rank 0's shard
    peak, peak_src = load_peaks()
    my_bytes = nb * BLOCK
    e_ms, d_ms = stage_ms["enc"], stage_ms["dec"]
    comp_share = (comp_bytes or 0) * nb / nblocks  # ~ this rank's compressed bytes
    # post-ZRLT length (m) of this rank's blocks: read from the block headers (mode byte + 3 length bytes)
    heads = d_blk[:, :4].cpu().numpy()
    m_total = int(sum(int.from_bytes(heads[i, 1:1 + 1 + ((heads[i, 0] >> 5) & 3)].tobytes(), "big")
                      for i in range(nb)))

    def rl(alg_bytes, ms):
        if not ms or ms <= 0:
            return None
        a = alg_bytes / (ms / 1e3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None,
                "algorithmic_bytes": int(alg_bytes), "ms": ms}

    stages = {
        "bwt_forward": rl(2 * my_bytes + 25 * nb, e_ms["bwt"]),
        "rank_forward": rl(2 * (my_bytes + 25 * nb), e_ms["rank"]),
        "zrlt_forward": rl(my_bytes + 25 * nb + m_total, e_ms["zrlt"]),
        "ans0_encode_kernel": rl(m_total + comp_share, e_ms["ans_enc_kernel"]),
        "ans0_encode_stage": rl(m_total + comp_share, e_ms["entropy"]),
        "ans0_decode_kernel": rl(m_total + comp_share, d_ms["ans_dec_kernel"]),
        "zrlt_inverse": rl(my_bytes + m_total, d_ms["zrlt"]),
        "rank_inverse": rl(2 * my_bytes, d_ms["rank"]),
        "bwt_inverse": rl(2 * my_bytes, d_ms["bwt"]),
        "encode_pipeline": rl(my_bytes + comp_share, e_ms["total"]),
        "decode_pipeline": rl(my_bytes + comp_share, d_ms["total"]),
    }
    # entropy kernels alone on the raw workload (BASELINE config 4 shape: -t NONE -e {ANS0,HUFFMAN}, 4 MiB blocks)
    try:
        for ename in ("ANS0", "HUFFMAN"):
            et2 = E_IDS[ename]
            tt0 = ctx.transform_type("NONE")
            sharded.encode_shard(ctx, tt0, et2, BLOCK, d_in, lens, BLOCK, d_blk, d_bits)
            sharded.encode_shard(ctx, tt0, et2, BLOCK, d_in, lens, BLOCK, d_blk, d_bits)
            t4 = ctx.timings()
            e4 = int((d_bits.sum().item() + 7) // 8)
            sharded.decode_shard(ctx, tt0, et2, BLOCK, d_blk, d_bits.cpu().numpy().astype(np.uint64), d_dec)
            t4d = ctx.timings()
            assert bool((d_dec == d_in).all().item())
            stages[f"{ename.lower()}_encode_kernel_config4"] = rl(my_bytes + e4, t4["ans_enc_kernel"])
            stages[f"{ename.lower()}_decode_kernel_config4"] = rl(my_bytes + e4, t4d["ans_dec_kernel"])
    except Exception as ex:
        stages["config4_error"] = str(ex)
    # DRAM traffic per launch from the committed `ncu --set full` captures (profiles/r01_ncu_*.md),
    # scaled to this rank's blocks: dram__bytes_read.sum + dram__bytes_write.sum
    NCU_TRAFFIC_PER_BLOCK = {
        # profiles/r01_ncu_traffic_64blocks_v6.md (stage sums of one encode+decode of 64 blocks) / 64
        "bwt_forward": 64.38e9 / 64,
        "rank_forward": 1.23e9 / 64,
        "zrlt_forward": 0.70e9 / 64,
        "ans0_encode_kernel": 0.46e9 / 64,
        "ans0_decode_kernel": 0.26e9 / 64,
        "zrlt_inverse": 1.07e9 / 64,
        "bwt_inverse": 24.17e9 / 64,
        "rank_inverse": (1.088619e9 + 1.040149e9) / 256,      # r01_ncu_rank_inverse_256blocks.md (--set full)
        "ans0_encode_kernel_config4": (491.811584e6 + 145.973248e6) / 64,  # r01_ncu_ans0_encode_v5_64blocks_config4.md
    }
    for k, per_block in NCU_TRAFFIC_PER_BLOCK.items():
        if stages.get(k):
            stages[k]["traffic"] = per_block * nb
    # the dominant KERNEL of the step: the inverse RANK chain (one launch, ~1/3 of the step on its own)
    dominant = dict(stages["rank_inverse"] or {})
    dominant["kernel"] = ("sbrt_inverse_fast_kernel<2> (inverse RANK: one dependency chain per block, one warp per "
                          "block; latency-bound, DRAM traffic = algorithmic bytes); the largest stage, BWT forward, "
                          "is ~100 launches and is listed under roofline_stages")
    dominant["peak_source"] = peak_src

    # ---- CPU baseline beside it: the unmodified reference, all host threads, on the FULL workload
    # (one pass, ~5-10 s): its stream hash is the bit-exactness check of the stream timed above.
    cores = os.cpu_count() or 1
    jobs = max(1, min(63, cores))
    cpu = None
    try:
        r = cpu_reference_run(data[:size], jobs)
        if r:
            cpu = {"value": size / (r["enc_s"] + r["dec_s"]) / 1e6, "unit": "MB/s", "cores": jobs,
                   "kind": "reference", "encode_MBps": size / r["enc_s"] / 1e6,
                   "decode_MBps": size / r["dec_s"] / 1e6,
                   "sample": f"the full {size >> 20} MiB workload, one pass, jobs={jobs}, in-memory streams"}
            info["reference_stream_sha256"] = r["sha256"]
            info["stream_matches_reference"] = bool(r["sha256"] == info.get("stream_sha256"))
            assert info["stream_matches_reference"], "GPU stream differs from the reference's stream"
    except AssertionError:
        raise
    except Exception as ex:  # the reference library did not travel: report the port instead
        cpu = {"error": str(ex)}
    if cpu is None:
        from oracle.oracle import Oracle
        o = Oracle()
        sample = 8 << 20
        t = time.perf_counter()
        c = o.stream_compress(data[:sample], TRANSFORM, ENTROPY, BLOCK)
        o.stream_decompress(c, sample)
        cpu = {"value": sample / (time.perf_counter() - t) / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
               "sample": "first 8 MiB of the workload, plain-C oracle, 1 thread"}

    line = {
        "metric": METRIC, "value": size / per_step / 1e6, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "encode_MBps": size * args.steps / enc_s / 1e6, "decode_MBps": size * args.steps / dec_s / 1e6,
        "config": {"workload": f"-t {TRANSFORM} -e {ENTROPY} -b 4m, {size >> 20} MiB synth_compressible(seed 2)",
                   "blocks": nblocks, "sharding": f"round-robin over {world} GPU(s)", "batch_blocks": batch,
                   "l2": "inputs (1 GiB) larger than L2; every step re-reads them from HBM",
                   "timing": "CUDA events on the library stream + wall clock, max over ranks"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": dominant,
        "roofline_stages": stages, "cpu_baseline": cpu, "stage_ms": {"encode": e_ms, "decode": d_ms},
    }
    line.update(info)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1 << 30)
    ap.add_argument("--batch", type=int, default=256, help="blocks per device batch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
