"""CPU tests of the CUDA kernels' logic: the same .cu sources compiled on top of the
execution-model emulator (tests/sim) are driven through the C ABI and compared with
the plain-C oracle.  (The real parity gate is tests/test_gpu_parity.py on the B200.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

import synth
from cases import small_cases, rng_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "libknzsim.so")


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    from kanzi_b200 import Context
    ctx = Context(0, 1 << 18, 4, lib_path=SIM)
    yield ctx
    ctx.close()


CASES = small_cases()


@pytest.mark.parametrize("ename", ["ANS0", "HUFFMAN", "ANS1", "FPAQ"])
def test_sim_entropy(sim, oracle, ename):
    for name, data in CASES.items():
        a, abits = sim.entropy_encode(ename, data)
        b, bbits = oracle.entropy_encode(ename, data)
        assert abits == bbits, (name, abits, bbits)
        assert np.array_equal(a, b), name
        dec = sim.entropy_decode(ename, b, bbits, data.size)
        assert np.array_equal(dec, data), name


@pytest.mark.parametrize("tname", ["ZRLT", "RANK", "MTFT", "BWT"])
def test_sim_stage_forward_inverse(sim, oracle, tname):
    chain = tname in ("RANK", "MTFT")  # one emulated warp per block walks the whole input: keep it short
    for name, data in CASES.items():
        n = data.size
        if chain and n > 20003 and name != "zero_heavy_70001":
            continue
        for cap in (n + 64, n):
            if tname == "BWT" and cap < n + 33:
                continue
            if chain and cap == n and n > 5000:
                continue
            a, applied = sim.transform_forward(tname, data, cap)
            b, flags = oracle.sequence_forward(tname, data, n, cap)
            assert applied == (flags != 0xFF), (name, tname, cap, applied, flags)
            if applied:
                assert a.size == b.size and np.array_equal(a, b), (name, tname, cap)
                back, ok = sim.transform_inverse(tname, b, n + 64)
                assert ok and np.array_equal(back, data), (name, tname, cap)


def test_sim_srt(sim, oracle):
    """SRT forward (tile tables + relabelled MTFT ranks + stable scatter) and the serial inverse."""
    for name, data in CASES.items():
        n = data.size
        if n > 70001:
            continue
        a, applied = sim.transform_forward("SRT", data, n + 1088)
        b, flags = oracle.sequence_forward("SRT", data, n + 1088, n + 1088)
        assert applied == (flags != 0xFF), (name, applied, flags)
        if applied:
            assert a.size == b.size and np.array_equal(a, b), name
            back, ok = sim.transform_inverse("SRT", b, n + 64)
            assert ok and np.array_equal(back, data), name
        _, applied = sim.transform_forward("SRT", data, n + 64)  # destination < n + 1024: refused
        assert not applied, name


@pytest.mark.parametrize("tname,ename", [("NONE", "ANS0"), ("BWT+RANK+ZRLT", "ANS0"), ("NONE", "NONE"),
                                         ("ZRLT", "ANS0"), ("BWT+MTFT+ZRLT", "ANS0"), ("BWT", "NONE"),
                                         ("NONE", "HUFFMAN"), ("BWT+RANK+ZRLT", "HUFFMAN"), ("NONE", "ANS1"),
                                         ("ZRLT", "ANS1"), ("BWT+SRT+ZRLT", "FPAQ")])
def test_sim_stream(sim, oracle, tname, ename):
    inputs = {
        "comp_150k": synth.synth_compressible(150000, 21),
        "text_70k": synth.synth_text(70000, 22),
        "incomp_80k": synth.synth_incompressible(80000, 23),
        "tiny_10": rng_bytes(10, 24),
        "tiny_16": rng_bytes(16, 25),
        "zeros_100k": np.zeros(100000, dtype=np.uint8),
        "tail_small": np.concatenate([synth.synth_text(65536, 26), rng_bytes(7, 27)]),
        "mixed": np.concatenate([synth.synth_text(65536, 26), synth.synth_incompressible(65536 + 13, 27)]),
    }
    chain = ("RANK" in tname) or ("MTFT" in tname)  # the emulated inverse chain is slow: shorter inputs, one block size
    for name, data in inputs.items():
        if chain:
            data = data[:40000] if name != "tail_small" else data
        for bs in ((16384,) if chain else (65536, 1 << 18)):
            a = sim.compress(data, tname, ename, bs)
            b = oracle.stream_compress(data, tname, ename, bs)
            assert a.size == b.size and np.array_equal(a, b), (name, tname, ename, bs, a.size, b.size)
            dec = sim.decompress(b, data.size)
            assert dec.size == data.size and np.array_equal(dec, data), (name, tname, ename, bs)


def test_sim_blocks(sim, oracle):
    bs = 65536
    data = synth.synth_compressible(3 * bs + 1000, 31)
    blocks = [data[i: i + bs] for i in range(0, data.size, bs)]
    enc = sim.encode_blocks(blocks, "BWT+RANK+ZRLT", "ANS0", bs)
    data_cap = max(bs + bs // 8, 262144)
    for i, blk in enumerate(blocks):
        ref_bytes, ref_bits = oracle.encode_block(blk, "BWT+RANK+ZRLT", "ANS0", data_cap, blocks[0].size + 33)
        assert enc[i][1] == ref_bits, (i, enc[i][1], ref_bits)
        assert np.array_equal(enc[i][0], ref_bytes), i
    dec = sim.decode_blocks([(e[0], e[1]) for e in enc], "BWT+RANK+ZRLT", "ANS0", bs)
    for i, blk in enumerate(blocks):
        assert np.array_equal(dec[i], blk), i


@pytest.mark.parametrize("tname,ename", [("LZ", "NONE"), ("LZX", "ANS0"), ("LZP", "NONE"), ("LZP+LZX", "HUFFMAN"),
                                         ("LZ+ZRLT", "ANS0")])
def test_sim_lz_family(sim, oracle, tname, ename):
    """LZ / LZX / LZP (csrc/lz.cu): the warp-per-block parse reproduces the reference's token, distance and
    length streams byte for byte; the decoders restore the input."""
    inputs = {
        "comp_200k": synth.synth_compressible(200000, 21),
        "text_70k": synth.synth_text(70000, 22),
        "incomp_80k": synth.synth_incompressible(80000, 23),          # stage refuses: skip flag
        "zeros_100k": np.zeros(100000, dtype=np.uint8),               # matches longer than 65535 + 254
        "period_1000": np.tile(rng_bytes(1000, 3), 150),
        "flag_bytes": np.tile(np.array([0xFC, 1, 2, 3, 0xFC, 0xFC, 7] * 40, dtype=np.uint8), 200),  # LZP escapes
        "twice_70k": np.concatenate([rng_bytes(70000, 9), rng_bytes(70000, 9), rng_bytes(5, 1)]),   # 3-byte distances
        "tiny_20": rng_bytes(20, 1),
        "text_100": synth.synth_text(100, 3),
    }
    for name, data in inputs.items():
        for bs in (65536, 1 << 18):
            want = oracle.stream_compress(data, tname, ename, bs)
            got = sim.compress(data, tname, ename, bs)
            assert got.size == want.size and np.array_equal(got, want), (name, tname, ename, bs, got.size, want.size)
            dec = sim.decompress(want, data.size)
            assert dec.size == data.size and np.array_equal(dec, data), (name, tname, ename, bs)
    # stage level: Transform<byte>::forward / inverse with the exact getMaxEncodedLength capacity
    first = tname.split("+")[0]
    data = inputs["comp_200k"][:65536]
    cap = data.size + data.size // 64 + (0 if first == "LZP" else 2)
    fwd, applied = sim.transform_forward(first, data, cap=cap)
    ref_out, flags = oracle.sequence_forward(first, data, data.size + 64, cap)
    assert applied == (flags != 0xFF) and (not applied or np.array_equal(fwd, ref_out))
    _, refused = sim.transform_forward(first, data, cap=cap - 1)  # one byte short: forward() returns false
    assert not refused
    if applied:
        back, ok = sim.transform_inverse(first, fwd, data.size + 4096)
        assert ok and np.array_equal(back, data)


def test_sim_ans1_both_decoders():
    """Order-1 rANS has two decode kernels: the shared-memory model (few long chunks, the default at these
    sizes) and the slot-table kernel (many chunks; forced here with KNZ_ANS1_SMEM_CHUNKS=0)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    code = r"""
import os, sys
import numpy as np
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "kanzi-cpp_b200"), os.path.join(ROOT, "tests")]
import synth
from cases import rng_bytes
from kanzi_b200 import Context
from oracle.oracle import Oracle
o = Oracle()
ctx = Context(0, 1 << 18, 4, lib_path=sys.argv[2])
for data in (synth.synth_compressible(300000, 21), synth.synth_text(70001, 22), rng_bytes(40, 5), np.zeros(5000, np.uint8),
             synth.synth_incompressible(70000, 3)):
    for bs in (65536, 1 << 18):
        want = o.stream_compress(data, "NONE", "ANS1", bs)
        assert np.array_equal(ctx.compress(data, "NONE", "ANS1", bs), want)
        assert np.array_equal(ctx.decompress(want, data.size), data), (data.size, bs)
print("ANS1_OK")
"""
    for limit in ("0", "600"):
        env = dict(os.environ, KNZ_ANS1_SMEM_CHUNKS=limit)
        out = subprocess.run([sys.executable, "-c", code, ROOT, SIM], env=env, capture_output=True, text=True, timeout=900)
        assert "ANS1_OK" in out.stdout, (limit, out.stdout[-500:], out.stderr[-1500:])


@pytest.mark.parametrize("ck", [32, 64])
def test_sim_block_checksums(oracle, ck):
    """XXHash32 / XXHash64 block checksums: written by the encoders (csrc/xxhash.cu + the block header kernel),
    verified on the device after the inverse transforms; a flipped payload bit is a CRC or bitstream error."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    from kanzi_b200 import Context, KanziGpuError
    ctx = Context(0, 1 << 16, 4, lib_path=SIM)
    inputs = [np.concatenate([synth.synth_compressible(150000, 21), synth.synth_incompressible(70000, 3),
                              np.zeros(10, np.uint8)]), rng_bytes(31, 5), rng_bytes(15, 6), synth.synth_text(65536 + 17, 7)]
    for data in inputs:
        for tname, ename in (("NONE", "ANS0"), ("BWT+RANK+ZRLT", "ANS0"), ("NONE", "NONE"), ("ZRLT", "HUFFMAN"),
                             ("NONE", "ANS1"), ("BWT+SRT+ZRLT", "FPAQ")):
            chain = ("RANK" in tname) or ("SRT" in tname)
            d = data[:40000] if chain else data
            bs = 16384 if chain else 65536
            want = oracle.stream_compress(d, tname, ename, bs, checksum=ck)
            ctx.set_checksum(ck)
            got = ctx.compress(d, tname, ename, bs)
            assert got.size == want.size and np.array_equal(got, want), (d.size, tname, ename, ck)
            ctx.set_checksum(0)  # decompress takes the checksum size from the stream header
            assert np.array_equal(ctx.decompress(want, d.size), d), (d.size, tname, ename, ck)
            if want.size > 1000:
                bad = want.copy()
                bad[bad.size - 40] ^= 0x04
                with pytest.raises(KanziGpuError) as ei:
                    ctx.decompress(bad, d.size)
                assert ei.value.code in (13, 15, 19), ei.value.code  # decode overflow, invalid bitstream or CRC
    # block-level entry points: the checksum follows the block length
    bs = 65536
    data = synth.synth_compressible(2 * bs + 100, 9)
    blocks = [data[i: i + bs] for i in range(0, data.size, bs)]
    ctx.set_checksum(ck)
    enc = ctx.encode_blocks(blocks, "ZRLT", "ANS0", bs)
    dec = ctx.decode_blocks([(e[0], e[1]) for e in enc], "ZRLT", "ANS0", bs)
    for i, blk in enumerate(blocks):
        assert np.array_equal(dec[i], blk), i
    hdr = 1 + 3  # mode byte + 3 length bytes for a 64 KiB block (len 65536 needs 3 bytes)
    stored = int.from_bytes(bytes(enc[0][0][hdr: hdr + ck // 8]), "big")
    assert stored == oracle.block_hash(blocks[0], ck)
    broken = enc[0][0].copy()
    broken[broken.size - 9] ^= 1
    with pytest.raises(KanziGpuError):
        ctx.decode_blocks([(broken, enc[0][1])], "ZRLT", "ANS0", bs)
    ctx.close()


def test_sim_decode_groups(oracle):
    """Block groups decoded on separate streams over disjoint workspace slices (knz_set_decode_groups)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    from kanzi_b200 import Context
    bs = 1 << 16
    ctx = Context(0, bs, 16, lib_path=SIM)
    data = synth.synth_compressible(9 * bs + 777, 5)
    comp = ctx.compress(data, "BWT+RANK+ZRLT", "ANS0", bs)
    assert np.array_equal(comp, oracle.stream_compress(data, "BWT+RANK+ZRLT", "ANS0", bs))
    for g in (2, 1):
        ctx.set_decode_groups(g)
        assert np.array_equal(ctx.decompress(comp, data.size), data), g
    ctx.close()


def test_sim_corrupt_streams():
    """Bit flips in a valid stream: the decoders must come back with an error or with (wrong) bytes,
    never hang or fault, and the context must stay usable (the block checksum option is off, as in
    BASELINE's configs, so silent corruption of literal bytes is expected)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
    from kanzi_b200 import Context, KanziGpuError
    bs = 1 << 16
    ctx = Context(0, bs, 8, lib_path=SIM)
    data = synth.synth_compressible(3 * bs + 500, 9)
    rng = np.random.RandomState(1)
    for tname, ename in (("NONE", "ANS0"), ("NONE", "HUFFMAN"), ("ZRLT", "ANS0"), ("LZ", "NONE"), ("LZP", "NONE"),
                         ("LZX", "HUFFMAN")):
        comp = ctx.compress(data, tname, ename, bs)
        outcomes = {"error": 0, "bytes": 0}
        for t in range(8):
            c = comp.copy()
            for _ in range(1 + t % 4):
                pos = rng.randint(24, c.size - 2)  # past the stream header
                c[pos] ^= 1 << rng.randint(0, 8)
            try:
                back = ctx.decompress(c, data.size)
                assert back.size <= data.size
                outcomes["bytes"] += 1
            except KanziGpuError:
                outcomes["error"] += 1
        assert outcomes["error"] + outcomes["bytes"] == 8
        assert np.array_equal(ctx.decompress(comp, data.size), data), (tname, ename, outcomes)
    ctx.close()


def _zrlt_inputs():
    """Inputs aimed at the mask forms of the ZRLT walks: zero runs of every length class inside a 16-byte
    segment, across segments and across 4 KiB tiles, dense 0xFE / 0xFF, ragged lengths."""
    rng = np.random.default_rng(20261017)
    out = {}
    for i, (n, pz, pbig) in enumerate([(16, 0.5, 0.1), (17, 0.5, 0.1), (31, 0.9, 0.0), (4096, 0.5, 0.02), (4097, 0.7, 0.2),
                                       (12301, 0.3, 0.5), (20011, 0.95, 0.01), (33333, 0.6, 0.05), (8192 + 15, 0.0, 1.0)]):
        a = rng.integers(1, 254, n, dtype=np.uint8)
        a[rng.random(n) < pbig] = rng.choice(np.array([0xFE, 0xFF], dtype=np.uint8), 1)[0]
        big = rng.random(n) < pbig
        a[big] = rng.choice(np.array([0xFE, 0xFF], dtype=np.uint8), int(big.sum()))
        a[rng.random(n) < pz] = 0
        out[f"rnd{i}_{n}"] = a
    # runs of exactly 1..40 zeros separated by one literal, then long runs that span tiles
    parts = []
    for L in list(range(1, 41)) + [62, 63, 64, 65, 126, 127, 128, 4095, 4096, 4097, 9000]:
        parts.append(np.zeros(L, dtype=np.uint8))
        parts.append(np.array([L & 0xFF or 7], dtype=np.uint8))
    out["run_ladder"] = np.concatenate(parts)
    out["run_ladder_shifted"] = np.concatenate([np.array([9, 0, 0], dtype=np.uint8), out["run_ladder"], np.zeros(21, dtype=np.uint8)])
    z = np.zeros(30000, dtype=np.uint8)
    z[::4099] = 0xFF
    out["sparse_ff"] = z
    return out


def test_sim_zrlt_masks(sim, oracle):
    for name, data in _zrlt_inputs().items():
        n = data.size
        cap = 2 * n + 64
        a, applied = sim.transform_forward("ZRLT", data, cap)
        b, flags = oracle.sequence_forward("ZRLT", data, n, cap)
        assert applied == (flags != 0xFF), (name, applied, flags)
        if applied:
            assert a.size == b.size and np.array_equal(a, b), name
            back, ok = sim.transform_inverse("ZRLT", b, n + 64)
            assert ok and np.array_equal(back, data), name


def test_sim_zrlt_inverse_token_classes(sim, oracle):
    """Inverse on arbitrary token streams (runs of 0xFF: lead / payload alternate; digits behind a lead are payloads)."""
    from kanzi_b200 import KanziGpuError
    rng = np.random.default_rng(7)
    alphabet = np.array([0, 1, 0xFF, 0xFF, 2, 3, 0x80, 0xFE], dtype=np.uint8)
    for n in (16, 33, 4096, 5000, 12345):
        for trial in range(3):
            src = alphabet[rng.integers(0, alphabet.size, n)]
            if trial == 2:
                src[rng.random(n) < 0.5] = 0xFF
            want, ok = oracle.sequence_inverse("ZRLT", 0, src, 1 << 17)
            try:
                got, applied = sim.transform_inverse("ZRLT", src, 1 << 17)
            except KanziGpuError:
                assert not ok, (n, trial)
                continue
            assert bool(ok) == applied, (n, trial)
            if ok:
                assert got.size == want.size and np.array_equal(got, want), (n, trial)


SRT_EDGE_SIZES = (40760, 4096 * 3 - 200, 4096 * 5 - 150)


def check_srt_header_across_tile_edge(ctx, oracle):
    """A ragged block whose SRT header (256 .. 1024 bytes) carries the length over a 4 KiB tile edge that the
    33-bytes-per-stage bound does not reach: the stage behind SRT must still see every byte (found by tools/fuzz_sim.py)."""
    for n in SRT_EDGE_SIZES:
        data = synth.synth_text(n, 1)
        for tname, ename in (("BWT+SRT+ZRLT", "ANS0"), ("SRT+ZRLT", "NONE")):
            got = ctx.compress(data, tname, ename, 65536)
            want = oracle.stream_compress(data, tname, ename, 65536)
            assert got.size == want.size and np.array_equal(got, want), (n, tname, ename)
            assert np.array_equal(ctx.decompress(got, n), data), (n, tname, ename)


def test_sim_srt_header_across_tile_edge(sim, oracle):
    check_srt_header_across_tile_edge(sim, oracle)
