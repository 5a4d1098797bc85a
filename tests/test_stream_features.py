"""Stream-layer features around the block path: `skipBlocks`, block ranges (`from` / `to`),
listener events.  CPU tests run the product's sources on the emulator (tests/sim) and compare with the
unmodified reference (oracle/_ref) driven through its Context constructors; the GPU tests do the same on
the B200 and, without the reference, fall back on size-independent properties."""
import os
import re
import subprocess

import numpy as np
import pytest

import synth
from cases import rng_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "libknzsim.so")


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    from kanzi_b200 import Context
    ctx = Context(0, 1 << 16, 4, lib_path=SIM)
    yield ctx
    ctx.close()


def mixed_input(bs, seed=11):
    """Blocks of different compressibility: text, random bytes, a block that starts with a ZIP signature,
    near-random (entropy just under the threshold is compressible: kept), a short tail."""
    parts = [synth.synth_text(bs, seed), rng_bytes(bs, seed + 1), synth.synth_text(bs, seed + 2),
             rng_bytes(bs, seed + 3, 200), rng_bytes(bs, seed + 4, 256), synth.synth_compressible(bs // 2 + 7, seed + 5)]
    parts[2][:4] = np.frombuffer(b"PK\x03\x04", dtype=np.uint8)
    return np.concatenate(parts)


def test_log2_table_matches_reference():
    """The entropy test's log2 table is generated (round(4096 * log2 i)); the reference ships it as a
    literal (Global.cpp:47-74).  All 257 entries must agree."""
    path = "/root/reference/src/Global.cpp"
    if not os.path.exists(path):
        pytest.skip("reference sources absent")
    import math
    m = re.search(r"LOG2_4096\[257\] = \{(.*?)\};", open(path).read(), re.S)
    ref = [int(x) for x in m.group(1).replace("\n", " ").split(",")]
    gen = [0] + [int(math.floor(4096.0 * math.log2(i) + 0.5)) for i in range(1, 257)]
    assert ref == gen


def check_skip_blocks(ctx, ref, bs, pipelines):
    data = mixed_input(bs)
    for tname, ename, ck in pipelines:
        ctx.set_checksum(ck)
        ctx.set_skip_blocks(True)
        got = ctx.compress(data, tname, ename, bs)
        ctx.set_skip_blocks(False)
        plain = ctx.compress(data, tname, ename, bs)
        ctx.set_checksum(0)
        want, _ = ref.stream_compress_ctx(data, tname, ename, bs, 1, ck, 1)
        assert got.size == want.size and np.array_equal(got, want), (tname, ename, ck)
        assert not np.array_equal(got, plain) or got.size != plain.size, "no block was skipped: the case tests nothing"
        back = ctx.decompress(got, data.size)
        assert np.array_equal(back, data), (tname, ename, ck)


def check_ranges(ctx, ref, bs):
    data = mixed_input(bs)
    nblk = (data.size + bs - 1) // bs
    ctx.set_skip_blocks(True)  # copy blocks (host path) and device blocks interleave
    comp = ctx.compress(data, "ZRLT", "ANS0", bs)
    ctx.set_skip_blocks(False)
    for lo, hi in [(1, 2), (2, 4), (3, nblk + 1), (nblk, nblk + 1), (1, nblk + 1), (4, 4 + 100), (nblk + 1, nblk + 3)]:
        got = ctx.decompress_range(comp, lo, hi, data.size)
        want = data[(lo - 1) * bs: min((hi - 1) * bs, data.size)]
        assert got.size == want.size and np.array_equal(got, want), (lo, hi)
        if ref is not None:
            r, rc, _ = ref.stream_decompress_ctx(comp, data.size, 1, lo, hi)
            assert rc == 0 and r.size == want.size and np.array_equal(r, want), ("reference", lo, hi)


def check_seek(ctx, bs):
    """Block offsets reported by the listener are valid seek positions (CompressedInputStream::seek)."""
    data = mixed_input(bs)
    evs = []
    ctx.set_listener(evs.append)
    comp = ctx.compress(data, "ZRLT", "ANS0", bs)
    ctx.set_listener(None)
    offs = {e["blockId"]: e["offset"] for e in evs if e["type"] == "BLOCK_INFO"}
    nblk = (data.size + bs - 1) // bs
    assert sorted(offs) == list(range(1, nblk + 1))
    for first, count in [(1, 1), (3, 2), (nblk, 1), (2, nblk)]:
        got = ctx.decompress_seek(comp, offs[first], count, data.size)
        want = data[(first - 1) * bs: min((first - 1 + count) * bs, data.size)]
        assert got.size == want.size and np.array_equal(got, want), (first, count)
    from kanzi_b200 import KanziGpuError
    with pytest.raises(KanziGpuError):
        ctx.decompress_seek(comp, offs[2] + 3, 1, data.size)  # not a block boundary


def norm_events(evs, decode):
    """Per block, the reference's order; BLOCK_INFO is compared apart (it is emitted at a different point of
    the interleaving in a threaded run, but with the same content)."""
    out = {}
    for e in evs:
        if e["type"] in ("BEFORE_TRANSFORM", "AFTER_TRANSFORM", "BEFORE_ENTROPY", "AFTER_ENTROPY", "BLOCK_INFO",
                         2, 3, 4, 5, 9):
            t = e["type"] if isinstance(e["type"], str) else {2: "BEFORE_TRANSFORM", 3: "AFTER_TRANSFORM",
                                                                4: "BEFORE_ENTROPY", 5: "AFTER_ENTROPY",
                                                                9: "BLOCK_INFO"}[e["type"]]
            hb = e.get("hashBits", {0: 0, 1: 32, 2: 64}.get(e.get("hashType", 0), 0))
            rec = (t, e["size"], (e["hash"] & 0xFFFFFFFFFFFFFFFF) if hb else 0, hb)
            if t == "BLOCK_INFO":
                rec = rec + (e["offset"], e["skipFlags"])
            out.setdefault(e["blockId"], []).append(rec)
    return out


def check_events(ctx, ref, bs):
    data = mixed_input(bs)
    for tname, ename, ck, skip in [("ZRLT", "ANS0", 32, 1), ("BWT+RANK+ZRLT", "HUFFMAN", 64, 0), ("NONE", "NONE", 0, 0)]:
        evs = []
        ctx.set_checksum(ck)
        ctx.set_skip_blocks(skip)
        ctx.set_listener(evs.append)
        comp = ctx.compress(data, tname, ename, bs)
        enc = norm_events(evs, False)
        evs.clear()
        back = ctx.decompress(comp, data.size)
        dec = norm_events(evs, True)
        ctx.set_listener(None)
        ctx.set_skip_blocks(False)
        ctx.set_checksum(0)
        assert np.array_equal(back, data)
        rcomp, revs = ref.stream_compress_ctx(data, tname, ename, bs, 1, ck, skip, events=True)
        assert np.array_equal(comp, rcomp)
        renc = norm_events(revs, False)
        assert enc == renc, (tname, ename, ck, skip)
        _, rc, rdevs = ref.stream_decompress_ctx(comp, data.size, 1, 0, 0, events=True)
        assert rc == 0
        assert dec == norm_events(rdevs, True), (tname, ename, ck, skip)


def test_sim_skip_blocks(sim, ref):
    check_skip_blocks(sim, ref, 1 << 16, [("ZRLT", "ANS0", 0), ("BWT+RANK+ZRLT", "ANS0", 32), ("NONE", "HUFFMAN", 64)])


def test_sim_block_ranges(sim, ref):
    check_ranges(sim, ref, 1 << 16)


def test_sim_events(sim, ref):
    check_events(sim, ref, 1 << 16)


def test_sim_seek(sim):
    check_seek(sim, 1 << 16)


# ---------------------------------------------------------------- on the B200
@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from kanzi_b200 import Context
    ctx = Context(0, 1 << 20, 16)
    yield ctx
    ctx.close()


def maybe_ref():
    from oracle.oracle import Ref
    return Ref.load()


@pytest.mark.gpu
def test_gpu_skip_blocks(gpu):
    r = maybe_ref()
    if r is None:
        pytest.skip("oracle/_ref/libkanzi_ref.so not present")
    check_skip_blocks(gpu, r, 1 << 20, [("BWT+RANK+ZRLT", "ANS0", 0), ("LZX", "HUFFMAN", 32), ("NONE", "ANS1", 64)])


@pytest.mark.gpu
def test_gpu_block_ranges(gpu):
    check_ranges(gpu, maybe_ref(), 1 << 20)


@pytest.mark.gpu
def test_gpu_seek(gpu):
    check_seek(gpu, 1 << 20)


@pytest.mark.gpu
def test_gpu_events(gpu):
    r = maybe_ref()
    if r is None:
        pytest.skip("oracle/_ref/libkanzi_ref.so not present")
    check_events(gpu, r, 1 << 20)
