"""CPU: the port's restatement of the engine's dynamic-mesh pass (oracle/imr_oracle.c imro_repose, IMR/shaders/dynamicMeshShader_glsl.comp:99-145)
is the checker of the device's re-pose kernel (tests/test_gpu_repose.py) -- the reference runs this pass in GLSL, so there is no CPU
reference to execute.  What pins the checker itself: an independent float64 restatement of the shader's formulas in numpy (agreement to
FP32 rounding), and cases whose answer is exact whatever the order of the sums (identity joints, a single joint of weight 1, no joints,
morph weights of zero)."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes


def _shader_f64(V, joints, weights, morph_w, M, IB):
    """:105-111 morph blend, :121-132 skin, in float64 and numpy's own summation order."""
    V = V.astype(np.float64)
    morphed = V[:, 0].copy()
    for i, w in enumerate(morph_w):
        morphed += float(w) * V[:, i + 1]
    if joints is None:
        return morphed
    M = M.astype(np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)          # column-major 16 floats -> row-major matrices
    IB = IB.astype(np.float64).reshape(-1, 4, 4).transpose(0, 2, 1)
    P = M @ IB
    out = np.zeros_like(morphed)
    n, G = joints.shape[0], joints.shape[1]
    for g in range(G):
        for c in range(4):
            out += weights[:, g, c, None].astype(np.float64) * np.einsum("nij,nj->ni", P[joints[:, g, c]], morphed)
    return out


@pytest.fixture(scope="module")
def ch():
    return scenes.character(nu=40, nv=20, n_joints=16, n_targets=2)


def test_port_repose_matches_an_independent_float64_restatement(port, ch):
    for phase in (0.0, 0.7, 2.9):
        M, mw = ch.pose(phase)
        got = port.repose(ch.vertices, 2, ch.joints, ch.weights, mw, M, ch.inverse_bind).astype(np.float64)
        want = _shader_f64(ch.vertices, ch.joints, ch.weights, mw, np.asarray(M), np.asarray(ch.inverse_bind))
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 4e-6 * scale, phase          # a few FP32 roundings of sums of ~20 terms
        assert np.allclose(got[:, 3], 1.0, atol=1e-5)                     # the weights sum to 1 and the matrices are affine: w stays 1


def test_port_repose_exact_cases(port, ch):
    n = ch.vertices.shape[0]
    J = ch.inverse_bind.shape[0] if hasattr(ch.inverse_bind, "shape") else len(ch.inverse_bind)
    eye = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (J, 1))
    zero_mw = np.zeros(2, np.float32)
    base = ch.vertices[:, 0].astype(np.float32)
    # no joints, no morphing: the base vertices, bit for bit
    assert np.array_equal(port.repose(ch.vertices, 2, None, None, zero_mw, None, None), base)
    # no joints: the morph blend alone, in the shader's order (base, then + w0 * t0, then + w1 * t1), one FP32 operation at a time
    mw = np.array([0.25, -0.5], np.float32)
    want = base.copy()
    for i in range(2):
        want = (want + (mw[i] * ch.vertices[:, i + 1].astype(np.float32)).astype(np.float32)).astype(np.float32)
    assert np.array_equal(port.repose(ch.vertices, 2, None, None, mw, None, None), want)
    # one joint of weight 1 per vertex, identity matrices: M * InvBind * v = v exactly
    jn = np.zeros((n, 1, 4), np.uint16); jn[:, 0, 0] = np.arange(n) % J
    w1 = np.zeros((n, 1, 4), np.float32); w1[:, 0, 0] = 1.0
    assert np.array_equal(port.repose(ch.vertices, 2, jn, w1, zero_mw, eye, eye), base)
    # one joint of weight 1, a pure translation: exactly base + t (w = 1), glm's (m0 v0 + m1 v1) + (m2 v2 + m3 v3) with two zero products
    T = eye.copy(); T[:, 12] = 2.0; T[:, 13] = -4.0; T[:, 14] = 0.5
    got = port.repose(ch.vertices, 2, jn, w1, zero_mw, T, eye)
    assert np.array_equal(got[:, :3], (base[:, :3] + np.array([2.0, -4.0, 0.5], np.float32)).astype(np.float32))
    # the product M * InvBind is what is applied: a matrix times its own inverse translation gives the identity back
    Tinv = eye.copy(); Tinv[:, 12] = -2.0; Tinv[:, 13] = 4.0; Tinv[:, 14] = -0.5
    assert np.array_equal(port.repose(ch.vertices, 2, jn, w1, zero_mw, T, Tinv), base)
