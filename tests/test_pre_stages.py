"""Host stages of the block pipeline: PACK / DNA (AliasCodec), MM (FSDCodec), UTF (UTFCodec),
stage by stage and inside whole streams, byte for byte against the unmodified reference (oracle/_ref).
CPU tests use the emulator build of the library (the host stages are the same code in both builds)."""
import os
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
    ctx = Context(0, 1 << 20, 4, lib_path=SIM)
    yield ctx
    ctx.close()


def utf8_text(n, seed, alphabets=("абвгдежзиклмнопрстуфхцчшщъыьэюя", "αβγδεζηθικλμνξοπρστυφχψω", "日本語中文字漢", "😀😁😂🤣")):
    r = np.random.RandomState(seed)
    words, size = [], 0
    while size < n + 16:
        a = alphabets[r.randint(0, len(alphabets))] if r.randint(0, 4) else "abcdefghij"
        w = "".join(a[r.randint(0, len(a))] for _ in range(r.randint(2, 9)))
        words.append(w)
        size += len(w.encode()) + 1
    b = (" ".join(words)).encode()
    return np.frombuffer(b[:n], dtype=np.uint8).copy()


def smooth16(n, seed):
    r = np.random.RandomState(seed)
    x = np.cumsum(r.randint(-40, 41, size=n // 2)).astype(np.int64) + 20000
    return (x & 0xFFFF).astype("<u2").view(np.uint8).copy()


def walk8(n, seed, step=4):
    r = np.random.RandomState(seed)
    out = np.zeros(n, dtype=np.uint8)
    for c in range(step):
        m = len(out[c::step])
        out[c::step] = (np.cumsum(r.randint(-3, 4, size=m)) + 128 + 17 * c) & 0xFF
    return out


def dna(n, seed):
    r = np.random.RandomState(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[r.randint(0, 4, size=n)].copy()


def stage_inputs():
    c = {}
    c["text_60000"] = synth.synth_text(60000, 3)
    c["text_5000"] = synth.synth_text(5000, 4)
    c["comp_200k"] = synth.synth_compressible(200000, 6)
    c["rnd4_9001"] = rng_bytes(9001, 5, 4) + 65
    c["rnd3_9002"] = rng_bytes(9002, 6, 3) + 48
    c["rnd16_9003"] = rng_bytes(9003, 7, 16) * 3
    c["rnd13_9000"] = rng_bytes(9000, 8, 13) + 100
    c["const_4000"] = np.full(4000, 0x41, dtype=np.uint8)
    c["two_sym_5001"] = rng_bytes(5001, 9, 2) * 255
    c["rnd200_30000"] = rng_bytes(30000, 10, 200)
    c["rnd256_30000"] = rng_bytes(30000, 11, 256)
    c["dna_40000"] = dna(40000, 12)
    c["dna_n_40001"] = np.where(rng_bytes(40001, 13, 50) == 0, ord("N"), dna(40001, 14)).astype(np.uint8)
    c["digits_20000"] = np.frombuffer(b"0123456789,.", dtype=np.uint8)[rng_bytes(20000, 15, 12)].copy()
    c["smooth16_100k"] = smooth16(100000, 16)
    c["walk8_step1"] = walk8(50000, 17, 1)
    c["walk8_step3"] = walk8(60000, 18, 3)
    c["walk8_step4"] = walk8(64000, 19, 4)
    c["walk8_step8"] = walk8(80000, 20, 8)
    c["utf8_50000"] = utf8_text(50000, 21)
    c["utf8_cyr_30000"] = utf8_text(30000, 22, ("абвгдежзиклмнопрстуфхцчшщъыьэюя",))
    c["utf8_bom"] = np.concatenate([np.frombuffer(b"\xef\xbb\xbf", dtype=np.uint8), utf8_text(20000, 23)])
    c["utf8_cut"] = utf8_text(40000, 24)[1:]  # may start in the middle of a sequence
    c["short_1000"] = synth.synth_text(1000, 25)
    bad = utf8_text(30000, 26)
    bad[15000] = 0xC0
    c["utf8_bad"] = bad
    return c


def english(n, seed, crlf=False, xml=False, odd=False):
    """Prose with capitalised sentence starts, static-dictionary words, repeated invented words, numbers;
    optionally CR+LF line ends, mark-up, and the bytes the text codec has to escape (0x0E, 0x0F, >= 0x80)."""
    r = np.random.RandomState(seed)
    common = ("the be and of in to with it that for you he have on said say at but we by had they as would who or can "
              "may do this was is much any from not she what their which people because through different between "
              "information everything government development organization").split()
    made = ["".join(chr(97 + r.randint(0, 26)) for _ in range(r.randint(3, 12))) for _ in range(400)]
    out, size, start = [], 0, True
    while size < n + 64:
        k = r.randint(0, 10)
        w = common[r.randint(0, len(common))] if k < 6 else made[int(r.randint(0, 400) ** 2 / 400)]
        if k == 9:
            w = str(r.randint(0, 100000))
        if start:
            w = w.capitalize()
        if xml and r.randint(0, 12) == 0:
            w = "<%s>%s&amp;</%s>" % (w, w, w)
        if odd and r.randint(0, 40) == 0:
            w += ["\x0f", "\x0e", "é", "ü\x0f"][r.randint(0, 4)]
        sep = " "
        start = False
        if r.randint(0, 12) == 0:
            sep = ". "
            start = True
        if r.randint(0, 30) == 0:
            sep = ("\r\n" if crlf else "\n")
        out.append(w + sep)
        size += len(w) + len(sep)
    b = "".join(out).encode("latin-1", "replace")
    return np.frombuffer(b[:n], dtype=np.uint8).copy()


def text_inputs():
    c = {}
    c["english_80000"] = english(80000, 41)
    c["english_crlf"] = english(50000, 42, crlf=True)
    c["english_xml"] = english(60000, 43, xml=True)
    c["english_odd"] = english(70000, 44, odd=True)
    c["english_all"] = english(120000, 45, crlf=True, xml=True, odd=True)
    c["spaces_first"] = np.concatenate([np.full(37, 32, dtype=np.uint8), english(30000, 46)])
    c["upper"] = np.frombuffer(bytes(english(40000, 47)).upper(), dtype=np.uint8).copy()
    c["grow"] = distinct_words(60000, 48)    # more distinct words than the initial list holds: it doubles
    c["wrap"] = distinct_words(700000, 49)   # more than 2^19: the oldest learnt words are replaced
    return c


def distinct_words(count, seed):
    r = np.random.RandomState(seed)
    letters = r.randint(0, 26, size=(count, 9)).astype(np.uint8) + 97
    letters[:, 8] = 32
    idx = np.arange(count)
    for k in range(4):  # four letters spell the word's number: all distinct
        letters[:, k] = 97 + (idx // (26 ** k)) % 26
    flat = letters.reshape(-1).copy()
    # repeat a slice so that learnt words are referenced again, old and recent ones
    return np.concatenate([flat, flat[: 9 * 5000], flat[-9 * 5000:]])


STAGES = ["PACK", "DNA", "MM", "UTF", "TEXT"]


@pytest.mark.parametrize("tname", STAGES)
def test_sim_pre_stage_vs_reference(sim, ref, tname):
    applied_some = False
    inputs = stage_inputs()
    if tname == "TEXT":
        inputs.update(text_inputs())
    for name, data in inputs.items():
        n = data.size
        cap = n + max(n // 16, 8192) + 1024  # >= getMaxEncodedLength of every stage, as EncodingTask sizes it
        want, flags, _ = ref.sequence_forward(tname, data, cap, cap)
        got, applied = sim.transform_forward(tname, data, cap)
        assert applied == (flags != 0xFF), (tname, name, applied, flags)
        if applied:
            applied_some = True
            assert got.size == want.size and np.array_equal(got, want), (tname, name)
            back, ok = sim.transform_inverse(tname, want, n + 64)
            assert ok and np.array_equal(back, data), (tname, name)
    assert applied_some, tname


def mixed_stream(bs, seed=31):
    """One block of each kind the host stages tell apart, and a short tail."""
    parts = [synth.synth_text(bs, seed), utf8_text(bs, seed + 1), smooth16(bs, seed + 2), dna(bs, seed + 3),
             rng_bytes(bs, seed + 4), (rng_bytes(bs, seed + 5, 3) + 48).astype(np.uint8), walk8(bs, seed + 6, 3),
             synth.synth_compressible(bs, seed + 7), np.full(bs, 7, dtype=np.uint8), english(bs, seed + 9),
             english(bs, seed + 10, crlf=True, xml=True, odd=True), utf8_text(bs // 3 + 11, seed + 8)]
    parts[4][:4] = np.frombuffer(b"RIFF", dtype=np.uint8)  # a container signature sets the data type up front
    return np.concatenate(parts)


PIPELINES = [("TEXT+UTF+PACK+MM+LZX", "HUFFMAN", 0),      # -l 3
             ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0", 32),      # -l 5
             ("TEXT+UTF+BWT+SRT+ZRLT", "FPAQ", 0),        # -l 6 (text codec variant 1)
             ("TEXT", "ANS1", 64), ("TEXT+LZ", "NONE", 0),
             ("PACK+LZX", "HUFFMAN", 0), ("DNA+LZ", "HUFFMAN", 0), ("MM+LZ", "ANS0", 32),
             ("UTF+PACK+MM+LZX", "HUFFMAN", 0), ("UTF+BWT+RANK+ZRLT", "ANS0", 64), ("PACK+MM", "NONE", 0),
             ("MM", "ANS1", 0)]


def check_streams(ctx, ref, bs, pipelines):
    data = mixed_stream(bs)
    for tname, ename, ck in pipelines:
        ctx.set_checksum(ck)
        got = ctx.compress(data, tname, ename, bs)
        ctx.set_checksum(0)
        want = ref.stream_compress(data, tname, ename, bs, 1, ck)
        assert got.size == want.size and np.array_equal(got, want), (tname, ename, ck)
        back = ctx.decompress(got, data.size)
        assert back.size == data.size and np.array_equal(back, data), (tname, ename, ck)
        r, rc = ref.stream_decompress(got, data.size)
        assert rc == 0 and np.array_equal(r, data), ("reference decoder", tname, ename)


def test_sim_pre_streams_vs_reference(sim, ref):
    # the emulator runs the device stages of every block on the CPU: the levels and one pipeline per host stage
    # (the GPU test runs the whole list at 1 MiB blocks)
    check_streams(sim, ref, 1 << 16, PIPELINES[:4] + [("DNA+LZ", "HUFFMAN", 0), ("UTF+PACK+MM+LZX", "HUFFMAN", 0), ("MM", "ANS1", 0)])


def test_sim_pre_stage_misplaced(sim):
    """A host stage behind a device stage is refused (the reference's levels put them first)."""
    from kanzi_b200 import KanziGpuError
    with pytest.raises(KanziGpuError):
        sim.compress(synth.synth_text(100000, 1), "LZX+PACK", "HUFFMAN", 1 << 16)


@pytest.mark.gpu
def test_gpu_pre_streams_vs_reference():
    import torch
    assert torch.cuda.is_available()
    from kanzi_b200 import Context
    from oracle.oracle import Ref
    r = Ref.load()
    if r is None:
        pytest.skip("oracle/_ref/libkanzi_ref.so not present")
    ctx = Context(0, 1 << 20, 16)
    try:
        check_streams(ctx, r, 1 << 20, PIPELINES)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_config3_levels_vs_reference():
    """BASELINE config 3 (`-l 3` = TEXT+UTF+PACK+MM+LZX / HUFFMAN, 4 MiB blocks, silesia-shaped input) and the
    level built on the headline pipeline (`-l 5` = TEXT+UTF+BWT+RANK+ZRLT / ANS0): streams identical to the
    reference's, round trip through both decoders."""
    import torch
    assert torch.cuda.is_available()
    from kanzi_b200 import Context
    from oracle.oracle import Ref
    r = Ref.load()
    if r is None:
        pytest.skip("oracle/_ref/libkanzi_ref.so not present")
    bs = 4 << 20
    data = synth.synth_silesia(48 * (1 << 20) + 12345, 3)
    ctx = Context(0, bs, 16)
    try:
        for tname, ename in [("TEXT+UTF+PACK+MM+LZX", "HUFFMAN"), ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0"), ("DNA+LZ", "HUFFMAN")]:
            got = ctx.compress(data, tname, ename, bs)
            want = r.stream_compress(data, tname, ename, bs, 1, 0)
            assert got.size == want.size and np.array_equal(got, want), (tname, ename)
            back = ctx.decompress(got, data.size)
            assert np.array_equal(back, data), (tname, ename)
    finally:
        ctx.close()
