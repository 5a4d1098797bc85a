"""The reference-side bindings executed: the reference's OWN stream classes (CompressedOutputStream /
CompressedInputStream, task pool, TransformSequence, block framing, bitstreams), compiled from the
unmodified sources with the three factory switches routed to integration/kanzi_gpu_adapters.hpp
(oracle/Makefile targets refgpu / refsim, integration/knz_reference_hooks.hpp), must produce the
very stream the unmodified reference produces, and decode it -- with several worker threads
sharing one context (the C ABI locks per context).

  CPU: the adapters on top of the emulator build of the library (libkanzi_refsim.so)
  GPU: the adapters on top of libknzgpu.so (libkanzi_refgpu.so, prebuilt, travels with the snapshot)
"""
import os
import subprocess

import numpy as np
import pytest

import synth
from cases import small_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _load(name):
    from oracle.oracle import Ref
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path) or not os.path.exists(os.path.join(REFDIR, "libkanzi_ref.so")):
        pytest.skip(f"oracle/_ref/{name} not built (reference sources absent)")
    return Ref(path), Ref(os.path.join(REFDIR, "libkanzi_ref.so"))


def _check_streams(gpuref, ref, inputs, pipelines, block_sizes, jobs):
    for name, data in inputs.items():
        for tname, ename in pipelines:
            for bs in block_sizes:
                # same job count on both sides: the reference's stream depends on it when a stage would expand a
                # block (the task buffers differ)
                want = ref.stream_compress(data, tname, ename, bs, jobs=jobs)
                got = gpuref.stream_compress(data, tname, ename, bs, jobs=jobs)
                assert got.size == want.size and np.array_equal(got, want), (name, tname, ename, bs)
                back, rc = gpuref.stream_decompress(want, data.size, jobs=jobs)
                assert rc == 0 and back.size == data.size and np.array_equal(back, data), (name, tname, ename, bs)


def test_adapters_on_emulator():
    if os.path.exists("/root/reference/src/Global.cpp"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "sim"), "-j8"], stdout=subprocess.DEVNULL)
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8", "refsim"],
                              stdout=subprocess.DEVNULL)
    gpuref, ref = _load("libkanzi_refsim.so")
    inputs = {"comp_90k": synth.synth_compressible(90000, 21), "text_40k": synth.synth_text(40000, 22),
              "tiny_10": small_cases()["rnd256_9"]}
    _check_streams(gpuref, ref, inputs, [("BWT+RANK+ZRLT", "ANS0"), ("NONE", "HUFFMAN"), ("BWT+SRT+ZRLT", "FPAQ"),
                                         ("ZRLT", "ANS1"), ("LZ", "HUFFMAN"), ("LZP+LZX", "ANS0")], (65536,), jobs=3)
    # stage level through the shim (TransformSequence built by the routed factory)
    data = inputs["text_40k"]
    for t in ("BWT", "RANK", "ZRLT", "BWT+RANK+ZRLT"):
        a, af, ok = gpuref.sequence_forward(t, data, data.size + 1200, data.size + 1200)
        b, bf, ok2 = ref.sequence_forward(t, data, data.size + 1200, data.size + 1200)
        assert af == bf and np.array_equal(a, b), t


@pytest.mark.gpu
def test_adapters_on_gpu():
    gpuref, ref = _load("libkanzi_refgpu.so")
    inputs = {"comp_3m": synth.synth_compressible(3 * (1 << 20) + 777, 31), "text_70k": synth.synth_text(70000, 22),
              "incomp_300k": synth.synth_incompressible(300000, 23), "tiny_10": small_cases()["rnd256_9"]}
    _check_streams(gpuref, ref, inputs, [("BWT+RANK+ZRLT", "ANS0"), ("BWT+MTFT+ZRLT", "HUFFMAN"), ("NONE", "ANS1"),
                                         ("BWT+SRT+ZRLT", "FPAQ"), ("ZRLT", "NONE"), ("LZ", "HUFFMAN"), ("LZX", "ANS0"),
                                         ("LZP", "NONE")], (65536, 1 << 20), jobs=8)
    big = synth.synth_compressible(20 << 20, 2)
    _check_streams(gpuref, ref, {"comp_20m": big}, [("BWT+RANK+ZRLT", "ANS0")], (4 << 20,), jobs=4)
