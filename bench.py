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


def entropy_sweep(dev, peak, total):
    """BASELINE config 4: -t NONE -e {HUFFMAN, ANS0, ANS1}, block 64 KiB .. 32 MiB, synth_compressible(seed 4),
    blocks resident in HBM.  Per cell: parity of the stream with the unmodified reference on a prefix, round
    trip of every block, and the entropy kernel alone (CUDA events inside the library) against the HBM
    roofline with m + e algorithmic bytes."""
    import ctypes
    import hashlib
    import torch
    import synth
    from kanzi_b200 import Context, E_IDS, _ptr
    from oracle.oracle import Ref
    ref = Ref.load()
    data = synth.synth_compressible(total, 4)
    d_all = torch.from_numpy(data).to(dev)
    out = {}
    for bs in (64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 32 << 20):
        nblocks = total // bs
        batch = min(nblocks, max(1, (256 << 20) // bs))
        ostride = (bs + bs // 4 + 4096 + 131072 * (bs // (4 << 20) + 1) + 255) // 256 * 256
        d_in = d_all[: nblocks * bs].view(nblocks, bs)
        for ename in ("HUFFMAN", "ANS0", "ANS1"):
            ctx = Context(dev.index, bs, batch)
            L = ctx.lib
            d_blk = torch.zeros((nblocks, ostride), dtype=torch.uint8, device=dev)
            d_bits = torch.zeros(nblocks, dtype=torch.int64, device=dev)
            d_dec = torch.zeros((nblocks, bs), dtype=torch.uint8, device=dev)
            lens = np.full(nblocks, bs, dtype=np.int32)
            ol = np.zeros(nblocks, dtype=np.int32)
            tt, et = ctx.transform_type("NONE"), E_IDS[ename]
            te = td = None
            rc_dec = 0
            for _ in range(2):  # second pass is the measured one
                rc = L.knz_encode_blocks_dev(ctx.h, tt, et, bs, d_in.data_ptr(), bs, _ptr(lens), nblocks, bs,
                                             d_blk.data_ptr(), ostride, d_bits.data_ptr(), None)
                assert rc == 0, (ename, bs, rc, L.knz_last_error(ctx.h))
                te = ctx.timings()
                hb = d_bits.cpu().numpy().astype(np.uint64)
                rc_dec = L.knz_decode_blocks_dev(ctx.h, tt, et, bs, d_blk.data_ptr(), ostride, _ptr(hb), nblocks,
                                                 d_dec.data_ptr(), bs, _ptr(ol))
                td = ctx.timings()
            same = (d_dec == d_in).all(dim=1).cpu().numpy()
            bad = [int(i) for i in np.nonzero(~same)[0]]
            e_bytes = int((int(d_bits.sum().item()) + 7) // 8)
            cell = {"block": bs, "blocks": nblocks, "ratio": e_bytes / (nblocks * bs)}
            # blocks the GPU decoder refuses are blocks the reference's own decoder refuses (normalisation
            # residual): check a few of them against it
            cell["undecodable_blocks"] = len(bad)
            if bad:
                assert ename == "ANS1" and rc_dec == 15, (ename, bs, rc_dec, bad[:4])
                if ref is not None:
                    for i in bad[:2]:
                        blk = data[i * bs:(i + 1) * bs]
                        enc, nbits = ref.entropy_encode(ename, blk)
                        dec, ok, _ = ref.entropy_decode(ename, enc, blk.size)
                        assert not (ok == blk.size and np.array_equal(dec, blk)), "reference decodes a block the GPU refuses"
                cell["undecodable_like_reference"] = True
            else:
                assert rc_dec == 0, (ename, bs, rc_dec)
            if ref is not None:  # stream parity on a prefix (whole blocks, at most 32 MiB)
                pre = data[: max(bs, min(total, 32 << 20) // bs * bs)]
                want = ref.stream_compress(pre, "NONE", ename, bs, jobs=min(16, os.cpu_count() or 1))
                got = ctx.compress(pre, "NONE", ename, bs)
                cell["stream_matches_reference"] = bool(got.size == want.size and np.array_equal(got, want))
                assert cell["stream_matches_reference"], (ename, bs)
            alg = nblocks * bs + e_bytes
            for leg, t in (("encode", te), ("decode", td)):
                k = t["ans_enc_kernel"] if leg == "encode" else t["ans_dec_kernel"]
                cell[leg] = {"kernel_ms": k, "stage_ms": t["entropy"],
                             "kernel_GBps": alg / (k / 1e3) / 1e9 if k > 0 else None,
                             "frac": alg / (k / 1e3) / 1e9 / peak if k > 0 else None,
                             "stage_GBps": alg / (t["entropy"] / 1e3) / 1e9 if t["entropy"] > 0 else None}
            out[f"{ename.lower()}_{bs >> 10}k"] = cell
            ctx.close()
            del d_blk, d_bits, d_dec
            torch.cuda.empty_cache()
    return out


def progress(rank, *a):
    """Progress notes on stderr (stdout carries the one JSON line)."""
    print(f"[bench rank {rank}] {time.strftime('%H:%M:%S')}", *a, file=sys.stderr, flush=True)


def run_ours(args, rank, world, local_rank):
    import ctypes
    import hashlib
    import torch
    import torch.distributed as dist
    import synth
    from kanzi_b200 import Context, E_IDS, _ptr

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    size = (args.size // BLOCK) * BLOCK
    nblocks = size // BLOCK
    data = synth.synth_compressible(size, 2)
    my = list(range(rank, nblocks, world))  # round-robin sharding (north_star): block i -> rank i % world
    nb = len(my)
    batch = max(1, min(args.batch, nb))
    ctx = Context(local_rank, BLOCK, batch)
    ctx.dist_init(rank, world)  # the library opens its own NCCL communicator (csrc/dist.cu)
    progress(rank, "context + communicator ready")
    L = ctx.lib
    host_in = torch.from_numpy(data).view(nblocks, BLOCK)[my].contiguous().pin_memory()
    d_in = torch.empty((max(nb, 1), BLOCK), dtype=torch.uint8, device=dev)
    d_in[:nb].copy_(host_in)
    d_dec = torch.empty((max(nb, 1), BLOCK), dtype=torch.uint8, device=dev)
    lens = np.full(max(nb, 1), BLOCK, dtype=np.int32)
    ttype = ctx.transform_type(TRANSFORM)
    etype = E_IDS[ENTROPY]
    hdr = np.zeros(32, dtype=np.uint8)
    hdr_bytes = L.knz_stream_header(ttype, etype, BLOCK, size, _ptr(hdr))
    stream_cap = (size + size // 4 + 65536 + 255) // 256 * 256
    d_stream = torch.zeros(stream_cap, dtype=torch.uint8, device=dev)  # rank 0 assembles; the others receive the broadcast
    all_bits = np.zeros(nblocks, dtype=np.uint64)
    out_lens = np.zeros(max(nb, 1), dtype=np.int32)
    stage_ms = {"enc": None, "dec": None}
    info = {}

    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {rc} {L.knz_last_error(ctx.h).decode()}")

    def encode_dev():
        """knz_dist_encode_dev: encode the rank's blocks, all-gather the bit counts, gather the payloads on
        rank 0 over NCCL, bit-concatenate the stream there (one C call)."""
        if rank == 0:
            d_stream.zero_()
        end = ctypes.c_uint64(0)
        check(L.knz_dist_encode_dev(ctx.h, ttype, etype, BLOCK, d_in.data_ptr(), BLOCK, _ptr(lens), nb, nblocks, BLOCK,
                                    d_stream.data_ptr(), stream_cap, 8 * hdr_bytes, _ptr(all_bits), ctypes.byref(end)),
              "knz_dist_encode_dev")
        stage_ms["enc"] = ctx.timings()
        return (int(end.value) + 8 + 7) // 8

    def decode_dev(nbytes):
        """knz_dist_decode_dev: broadcast the assembled stream from rank 0, decode the rank's blocks out of it."""
        check(L.knz_dist_decode_dev(ctx.h, ttype, etype, BLOCK, d_stream.data_ptr(), (nbytes + 255) // 256 * 256,
                                    8 * hdr_bytes, _ptr(all_bits), nblocks, d_dec.data_ptr(), BLOCK, _ptr(out_lens)),
              "knz_dist_decode_dev")
        assert (out_lens[:nb] == BLOCK).all()
        stage_ms["dec"] = ctx.timings()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        nbytes = encode_dev()
        if world > 1:  # every rank needs the stream length (8 bytes; the stream itself travels inside the C call)
            t = torch.tensor([nbytes], dtype=torch.int64, device=dev)
            dist.broadcast(t, 0)
            nbytes = int(t.item())
        t_mid = time.perf_counter()
        decode_dev(nbytes)
        return nbytes, t_mid

    # ---- warm-up + correctness of what is being timed
    comp_bytes = None
    for _ in range(max(1, args.warmup)):
        comp_bytes, _ = step_dev()
    barrier()
    progress(rank, "warm-up done")
    assert bool((d_dec[:nb] == d_in[:nb]).all().item()), "decode(encode(x)) != x"
    if rank == 0:
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
    progress(rank, f"device-resident leg done: {per_step * 1e3:.1f} ms/step")

    # ---- e2e through the public API with HOST buffers (pinned): knz_compress / knz_decompress at N = 1,
    # knz_compress_dist / knz_decompress_dist at N > 1 -- every rank holds the input (compress) and the
    # stream (decompress) in host memory, as N processes reading the same file would; rank 0 receives the
    # stream, every rank receives the blocks it owns.  Host -> device and device -> host copies are inside
    # the timed region; handing the stream from the compress leg to the other ranks' hosts (a file, in real
    # use) is outside it.
    host_full = torch.from_numpy(data).pin_memory().numpy()
    out_comp = torch.empty(stream_cap if rank == 0 else 16, dtype=torch.uint8).pin_memory().numpy()
    out_plain = torch.empty(size, dtype=torch.uint8).pin_memory().numpy()
    comp_host = torch.empty(stream_cap, dtype=torch.uint8).pin_memory()

    def e2e_compress():
        if world == 1:
            return ctx.compress(host_full, TRANSFORM, ENTROPY, BLOCK, out=out_comp)
        return ctx.compress_dist(host_full, TRANSFORM, ENTROPY, BLOCK, out=out_comp)

    def share_stream(comp):
        """Untimed: the stream written by rank 0 reaches the other ranks' host memory (the file they would read)."""
        n = torch.tensor([comp.size], dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(n, 0)
        n = int(n.item())
        if rank == 0:
            comp_host[:n].copy_(torch.from_numpy(comp))
        if world > 1:
            t = comp_host[:n].to(dev)
            dist.broadcast(t, 0)
            comp_host[:n].copy_(t.cpu())
        return comp_host[:n].numpy()

    def e2e_decompress(comp):
        if world == 1:
            return ctx.decompress(comp, size, out=out_plain)
        return ctx.decompress_dist(comp, size, out=out_plain)

    comp = share_stream(e2e_compress())  # warm
    e2e_decompress(comp)
    barrier()
    e_enc = e_dec = 0.0
    for _ in range(args.steps):
        barrier()
        ts = time.perf_counter()
        c = e2e_compress()
        barrier()
        e_enc += time.perf_counter() - ts
        comp = share_stream(c)
        barrier()
        ts = time.perf_counter()
        back = e2e_decompress(comp)
        barrier()
        e_dec += time.perf_counter() - ts
    assert back.size == size
    for i in my[:4] + my[-4:]:
        assert np.array_equal(out_plain[i * BLOCK:(i + 1) * BLOCK], data[i * BLOCK:(i + 1) * BLOCK]), "e2e round trip"
    if rank == 0:
        assert hashlib.sha256(comp.tobytes()).hexdigest() == info["stream_sha256"], "e2e stream != device-resident stream"
    t_e = torch.tensor([e_enc, e_dec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e_enc, e_dec = [float(x) / args.steps for x in t_e.tolist()]
    progress(rank, f"e2e leg done: {(e_enc + e_dec) * 1e3:.1f} ms/step")
    e2e = {"value": size / (e_enc + e_dec) / 1e6, "unit": "MB/s",
           "h2d_bytes_per_step": int(size + comp.size), "d2h_bytes_per_step": int(comp.size + size),
           "encode_MBps": size / e_enc / 1e6, "decode_MBps": size / e_dec / 1e6,
           "api": ("knz_compress + knz_decompress" if world == 1 else "knz_compress_dist + knz_decompress_dist")
                  + " (host pinned buffers; at N > 1 every rank uploads its own blocks / its own blocks' bit ranges)"}

    # the library's communicator is torn down while every rank is still here (communicator teardown
    # synchronises the ranks; interleaving it with torch's own teardown on another rank can deadlock)
    launch_count_total = launches
    barrier()
    ctx.close()
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (stage level, SURVEY.md §8(d) algorithmic bytes), rank 0's shard
    peak, peak_src = load_peaks()
    my_bytes = nb * BLOCK
    e_ms, d_ms = stage_ms["enc"], stage_ms["dec"]
    comp_share = (comp_bytes or 0) * nb / nblocks  # ~ this rank's compressed bytes
    # post-ZRLT length (m) of this rank's blocks: read from the block headers inside the assembled stream
    # (5 + lw bit prefix, then mode byte + 1..4 length bytes)
    m_total, pos = 0, 8 * hdr_bytes
    sbits = np.unpackbits(stream[: min(stream.size, comp_bytes)]) if nblocks <= 4096 else None
    for i in range(nblocks):
        w = int(all_bits[i])
        lw = 3
        while lw < 35 and (w >> lw) != 0:
            lw += 1
        start = pos + 5 + lw
        if i % world == rank and sbits is not None:
            hb = np.packbits(sbits[start: start + 40]).tobytes()
            ds = 1 + ((hb[0] >> 5) & 3)
            o = 2 if (hb[0] & 0x10) else 1
            m_total += int.from_bytes(hb[o: o + ds], "big")
        pos = start + w

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
    # BASELINE config 4: -t NONE -e {HUFFMAN, ANS0, ANS1} x block size, device-resident, kernel-only GB/s
    del d_in, d_dec, d_stream
    torch.cuda.empty_cache()
    sweep = {}
    if world == 1 and not args.no_sweep:
        try:
            sweep = entropy_sweep(dev, peak, min(size, args.sweep_mib << 20))
            for k, v in sweep.items():
                stages["config4_" + k] = v
        except Exception as ex:
            stages["config4_error"] = repr(ex)
    # DRAM traffic from the committed ncu captures, scaled to this rank's blocks:
    # dram__bytes_read.sum + dram__bytes_write.sum
    NCU_TRAFFIC_PER_BLOCK = {
        # profiles/r02_traffic_64blocks.md (stage sums of one encode+decode pass over 64 blocks) / 64
        "bwt_forward": 64.568e9 / 64,
        "rank_forward": 1.230e9 / 64,
        "zrlt_forward": 0.721e9 / 64,                         # profiles/r02e_zrlt_launches.md (mask walks, staged stores)
        "ans0_encode_kernel": 0.458e9 / 64,
        "ans0_decode_kernel": 0.248e9 / 64,
        "zrlt_inverse": 0.825e9 / 64,                         # profiles/r02e_zrlt_launches.md
        "bwt_inverse": 24.261e9 / 64,
        "rank_inverse": (1.096665e9 + 1.040414e9) / 256,      # profiles/r02_ncu_rank_256blocks.md (--set full)
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
                   "timing": "CUDA events on the library stream + wall clock, max over ranks",
                   "stream_equivalence": "byte-identical to the reference with any -j on this workload; where ZRLT would "
                                         "expand a block the reference's stream depends on -j and the GPU path "
                                         "reproduces -j 1 (DESIGN.md, known reference quirks 1)"},
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
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-4 entropy sweep (N = 1 only)")
    ap.add_argument("--sweep-mib", type=int, default=256, help="bytes per cell of the config-4 sweep")
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
