"""GPU, BASELINE config 5: re-posing on the device (imrcd_skin_create / imrcd_mesh_bind_skin / imrcd_meshes_repose): the arithmetic of the
engine's dynamic-mesh compute pass (dynamicMeshShader_glsl.comp:99-145).  The reference runs it as GLSL, so the checker is the port's
restatement (oracle imro_repose): re-posed vertices bit for bit, the tree's triangles = those vertices through the mesh's vertex ids,
then refit + collide against the port on the exported trees; at 256 characters x 20,164 triangles in one batched pass."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import (CollisionDetection, ImrcdError, OBBtree, Skin, last_repose_ms, refit_meshes, repose_meshes,
                                            reposed_vertices)
from helpers import compare_frame, f32_bits, gpu_frame, oracle_frame
from test_gpu_build import check_boxes_contain, check_tree_structure

pytestmark = pytest.mark.gpu


def _expected(port, ch, mats, mw):
    return port.repose(ch.vertices, ch.vertices.shape[1] - 1, ch.joints, ch.weights, mw, mats, ch.inverse_bind)


def test_repose_bit_exact_and_refit(gpu_ctx, port):
    ch = scenes.character(40, 20, n_joints=16, n_targets=2)
    skin = Skin(gpu_ctx, ch.vertices, ch.joints, ch.weights)
    trees = [OBBtree(gpu_ctx, ch.mesh.positions, ch.mesh.normals, ch.mesh.vertex_ids) for _ in range(3)]
    for t in trees:
        t.bind_skin(skin)
    poses = [ch.pose(0.4 + 1.3 * k) for k in range(3)]
    repose_meshes(gpu_ctx, trees, np.stack([p[1] for p in poses]), np.stack([p[0] for p in poses]), np.stack([ch.inverse_bind] * 3))
    got = reposed_vertices(gpu_ctx, 3 * skin.n_vertices).reshape(3, -1, 4)
    refit_meshes(gpu_ctx)
    for k, t in enumerate(trees):
        want = _expected(port, ch, poses[k][0], poses[k][1])
        assert np.array_equal(f32_bits(got[k]), f32_bits(want))
        assert np.abs(want[:, :3] - ch.vertices[:, 0, :3]).max() > 0.01          # the pose moved something
        posed = scenes.Mesh(np.ascontiguousarray(want[:, :3][ch.mesh.vertex_ids].reshape(-1, 9)), ch.mesh.normals, ch.mesh.vertex_ids)
        flat = t.export()
        lo, hi = check_tree_structure(flat, posed)            # the tree's triangles are the re-posed vertices through the vertex ids
        check_boxes_contain(flat, lo, hi)
    assert last_repose_ms(gpu_ctx) > 0.0


def test_repose_morph_only_and_identity(gpu_ctx, port):
    ch = scenes.character(24, 12, n_joints=8, n_targets=3)
    morph = Skin(gpu_ctx, ch.vertices)                       # no joints: result = morphed vertex (:121-123)
    tree = OBBtree(gpu_ctx, ch.mesh.positions, ch.mesh.normals, ch.mesh.vertex_ids)
    tree.bind_skin(morph)
    mw = np.array([[0.3, -0.7, 1.1]], np.float32)
    repose_meshes(gpu_ctx, [tree], mw)
    got = reposed_vertices(gpu_ctx, morph.n_vertices)
    assert np.array_equal(f32_bits(got), f32_bits(port.repose(ch.vertices, 3, None, None, mw[0], None, None)))
    # identity joints and zero morph weights give the bind pose back exactly
    skin = Skin(gpu_ctx, ch.vertices, ch.joints, ch.weights)
    tree.bind_skin(skin)
    eye = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (1, ch.n_joints, 1))
    repose_meshes(gpu_ctx, [tree], np.zeros((1, 3), np.float32), eye, eye)
    got = reposed_vertices(gpu_ctx, skin.n_vertices)
    assert np.allclose(got[:, :3], ch.vertices[:, 0, :3], atol=2e-6)


def test_repose_argument_checks(gpu_ctx):
    ch = scenes.character(16, 8, n_joints=8, n_targets=1)
    skin = Skin(gpu_ctx, ch.vertices, ch.joints, ch.weights)
    tree = OBBtree(gpu_ctx, ch.mesh.positions, ch.mesh.normals, ch.mesh.vertex_ids)
    with pytest.raises(ImrcdError):
        repose_meshes(gpu_ctx, [tree], np.zeros((1, 1), np.float32))               # not bound to a skin
    small = Skin(gpu_ctx, ch.vertices[:10])
    with pytest.raises(ImrcdError):
        tree.bind_skin(small)                                                       # vertex ids beyond the skin's vertices
    tree.bind_skin(skin)
    eye = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (1, 4, 1))
    with pytest.raises(ImrcdError):
        repose_meshes(gpu_ctx, [tree], np.zeros((1, 1), np.float32), eye, eye)     # the skin names joints 0..7, four matrices given


def test_config5_full_size_repose_refit_collide(gpu_ctx, port):
    """256 characters x 20,164 triangles re-posed in one pass, refitted, collided: vertices of sampled characters against the port, every
    sampled box against its triangles, the frame against the port on the exported trees for sampled pairs."""
    ch = scenes.character()
    assert ch.mesh.n_tri == 20164
    n = 256
    skin = Skin(gpu_ctx, ch.vertices, ch.joints, ch.weights)
    trees = [OBBtree(gpu_ctx, ch.mesh.positions, ch.mesh.normals, ch.mesh.vertex_ids) for _ in range(n)]
    for t in trees:
        t.bind_skin(skin)
    poses = [ch.pose(0.37 * k) for k in range(n)]
    jm = np.stack([p[0] for p in poses]); mw = np.stack([p[1] for p in poses]); ib = np.stack([ch.inverse_bind] * n)
    for _ in range(2):                                         # twice: the second pass runs on the cached fit lists
        repose_meshes(gpu_ctx, trees, mw, jm, ib)
        refit_ms = refit_meshes(gpu_ctx)
    got = reposed_vertices(gpu_ctx, n * skin.n_vertices).reshape(n, -1, 4)
    for k in (0, 97, 255):
        want = _expected(port, ch, poses[k][0], poses[k][1])
        assert np.array_equal(f32_bits(got[k]), f32_bits(want))
        posed = scenes.Mesh(np.ascontiguousarray(want[:, :3][ch.mesh.vertex_ids].reshape(-1, 9)), ch.mesh.normals, ch.mesh.vertex_ids)
        flat = trees[k].export()
        lo, hi = check_tree_structure(flat, posed)
        check_boxes_contain(flat, lo, hi, sample=200)
    assert 0.0 < refit_ms < 50.0
    sc = scenes.scene_instances(ch.mesh, n, seed=7, neighbours=6.0)
    cd = CollisionDetection(ctx=gpu_ctx)
    ids = np.array([t.mesh_id for t in trees], np.uint32)
    cd.Reset(); cd.add_entries(sc.matrices, ids, sc.should_callback, sc.entities); cd.ExecuteCollisionDetection()
    st = cd.stats(); bp = cd.broad_pairs(); ep, hits = cd.results(want_hits=True)
    assert st["n_hits"] > 1000 and st["n_colliding"] > 10
    order = np.argsort(hits["pair"], kind="stable"); hp = hits["pair"][order]
    rng = np.random.default_rng(3)
    p_trees = {}
    for k in rng.choice(np.unique(hp), 12, replace=False).tolist():
        i, j = bp[k].tolist()
        for e in (i, j):
            if e not in p_trees:
                p_trees[e] = port.tree_import(trees[e].export())
        r = port.pair(p_trees[i], sc.matrices[i], p_trees[j], sc.matrices[j])
        lo, hi = np.searchsorted(hp, k), np.searchsorted(hp, k, side="right")
        gh = hits[order[lo:hi]]
        o_rec = {(int(a), int(b)): f32_bits(seg).tobytes() for (a, b), seg in zip(r.hit_ids.tolist(), r.hit_seg)}
        g_rec = {(int(h["tri_first"]), int(h["tri_second"])): f32_bits(np.concatenate([h["source"], h["target"], [h["weight"]]])).tobytes() for h in gh}
        assert o_rec == g_rec, k
