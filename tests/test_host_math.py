"""The arithmetic header of the CUDA kernels (csrc/lb200_math.cuh) compiled for the HOST against the reference-run vectors: every
function must reproduce the reference's bits, which pins the operation order the device code is written in (the device build routes the
same expressions through __fmul_rn / __fadd_rn / ... so that no compiler flag can fuse them).  No GPU needed."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def hm():
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.isdir(cuda_inc):
        pytest.skip("CUDA headers not found")
    tmp = tempfile.mkdtemp()
    so = os.path.join(tmp, "libhost_math.so")
    # the flags the reference is built with: SSE2, no FMA contraction (scripts/genie.lua:339-342)
    cmd = ["/usr/bin/g++", "-x", "c++", "-std=c++17", "-O2", "-msse2", "-ffp-contract=off", "-fPIC", "-shared", "-I", cuda_inc,
           "-I", os.path.join(ROOT, "lumixengine_b200", "csrc"), os.path.join(ROOT, "tests", "host_math", "host_math_harness.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(so)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_device_math_header_reproduces_reference_bits(hm):
    k = np.load(os.path.join(G, "math_kat.npz"))
    n = len(k["qa"])
    c = lambda a: np.ascontiguousarray(a)  # noqa: E731
    o4, o3 = np.zeros((n, 4), np.float32), np.zeros((n, 3), np.float32)
    hm.hm_quat_mul(P(c(k["qa"])), P(c(k["qb"])), P(o4), C.c_uint(n)); assert np.array_equal(bits(o4), bits(k["quat_mul"]))
    hm.hm_quat_rotate(P(c(k["qa"])), P(c(k["v"])), P(o3), C.c_uint(n)); assert np.array_equal(bits(o3), bits(k["quat_rotate"]))
    hm.hm_simd_nlerp(P(c(k["qa"])), P(c(k["qb"])), P(c(k["t"])), P(o4), C.c_uint(n)); assert np.array_equal(bits(o4), bits(k["simd_nlerp"]))
    hm.hm_lerp(P(c(k["v"])), P(c(k["v2"])), P(c(k["t"])), P(o3), C.c_uint(n)); assert np.array_equal(bits(o3), bits(k["lerp"]))
    o7, o8, o16 = np.zeros((n, 7), np.float32), np.zeros((n, 8), np.float32), np.zeros((n, 16), np.float32)
    hm.hm_lrt_mul(P(c(k["lrt_a"])), P(c(k["lrt_b"])), P(o7), C.c_uint(n)); assert np.array_equal(bits(o7), bits(k["lrt_mul"]))
    hm.hm_lrt_to_dual_quat(P(c(k["lrt_a"])), P(o8), C.c_uint(n)); assert np.array_equal(bits(o8), bits(k["lrt_to_dual_quat"]))
    hm.hm_lrt_to_matrix(P(c(k["lrt_a"])), P(o16), C.c_uint(n)); assert np.array_equal(bits(o16), bits(k["lrt_to_matrix"]))
    out = np.zeros((n, 56), np.uint8)
    hm.hm_transform_compose(P(c(k["tr_a"])), P(c(k["tr_b"])), P(out), C.c_uint(n))
    assert np.array_equal(out[:, :52], k["compose"][:, :52])


def test_bone_attachment_expression(hm):
    """The expression bone_attachments_kernel evaluates, on the host, against the reference-run vectors of updateBoneAttachment."""
    k = np.load(os.path.join(G, "world_kat.npz"))
    n = len(k["ba_parent"])
    out = np.zeros((n, 56), np.uint8)
    c = np.ascontiguousarray
    hm.hm_bone_attachments(P(c(k["ba_parent"])), P(c(k["ba_bone"])), P(c(k["ba_rel"])), P(c(k["ba_scale"])), P(out), C.c_uint(n))
    assert np.array_equal(out[:, :52], k["ba_out"][:, :52])
