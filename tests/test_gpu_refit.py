"""GPU, BASELINE config 5: re-posed meshes -- update positions, refit the OBB tree (same topology), collide.
The reference has no refit (SURVEY finding 4); parity is asserted (a) bit-exactly against the port traversing the exported
refitted tree and (b) on tree-independent outputs against the reference REBUILDING its tree from the re-posed triangles."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import IMRCD_BUILD_REFERENCE, CollisionDetection, OBBtree, refit_meshes
from helpers import compare_frame, gpu_frame, oracle_frame
from test_gpu_build import _assert_tree_independent_parity, check_boxes_contain, check_tree_structure

pytestmark = pytest.mark.gpu


def repose(mesh, phase):
    """A smooth non-rigid deformation (twist about z + bend), the kind of motion skinning produces."""
    p = mesh.positions.reshape(-1, 3).astype(np.float64)
    ang = 0.6 * np.sin(phase) * p[:, 2] + 0.3 * np.cos(2 * phase) * p[:, 0]
    c, s = np.cos(ang), np.sin(ang)
    q = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1], p[:, 2] + 0.25 * np.sin(phase + 2.0 * p[:, 0])], 1)
    return scenes.Mesh(np.ascontiguousarray(q.astype(np.float32).reshape(-1, 9)), mesh.normals, mesh.vertex_ids, mesh.name + f"@{phase:.2f}")


@pytest.mark.parametrize("mode", [0, IMRCD_BUILD_REFERENCE], ids=["morton", "reference-built"])
def test_refit_structure_and_containment(gpu_ctx, mode):
    base = scenes.torus(60, 30)
    tree = OBBtree(gpu_ctx, base.positions, base.normals, base.vertex_ids, build_mode=mode)
    before = tree.export()
    for phase in (0.7, 2.1):
        posed = repose(base, phase)
        tree.update_positions(posed.positions)
        tree.refit()
        flat = tree.export()
        assert np.array_equal(flat.tri_orig, before.tri_orig) and np.array_equal(flat.left, before.left) and np.array_equal(flat.tri_off, before.tri_off)
        lo, hi = check_tree_structure(flat, posed)          # leaf order unchanged, triangles replaced by the re-posed ones
        check_boxes_contain(flat, lo, hi)


def test_refit_many_meshes_in_one_pass(gpu_ctx):
    base = [scenes.torus(30, 14), scenes.uv_sphere(20, 15), scenes.box_mesh(1, 1, 1, sub=1), scenes.cylinder(16, 6)]
    tiny = scenes.Mesh(base[2].positions[:3].copy(), base[2].normals[:3].copy(), base[2].vertex_ids[:3].copy(), "tiny3")   # leaf root
    base.append(tiny)
    a = [OBBtree(gpu_ctx, m.positions, m.normals, m.vertex_ids) for m in base]
    b = [OBBtree(gpu_ctx, m.positions, m.normals, m.vertex_ids) for m in base]
    posed = [repose(m, 1.3 + k) for k, m in enumerate(base)]
    for t, m in zip(a, posed):
        t.update_positions(m.positions); t.refit()                      # one by one
    for t, m in zip(b, posed):
        t.update_positions(m.positions)
    assert refit_meshes(gpu_ctx) >= 0.0                                  # all dirty meshes, one batched pass
    for ta, tb, m in zip(a, b, posed):
        fa, fb = ta.export(), tb.export()
        assert np.array_equal(fa.boxes.view(np.uint32), fb.boxes.view(np.uint32))
        lo, hi = check_tree_structure(fb, m)
        check_boxes_contain(fb, lo, hi)


def test_frame_after_refit(gpu_ctx, port, oracle):
    """256-character style frame in miniature: several re-posed variants of one mesh, many instances, collide."""
    base = scenes.torus(60, 30)
    variants = [repose(base, 0.9 * k) for k in range(4)]
    trees = [OBBtree(gpu_ctx, base.positions, base.normals, base.vertex_ids) for _ in variants]      # all built in the rest pose
    for t, m in zip(trees, variants):
        t.update_positions(m.positions)
    refit_meshes(gpu_ctx, trees)
    n = 160
    sc = scenes.scene_instances(base, n, seed=41, neighbours=7.0)
    scene = scenes.Scene(variants, (np.arange(n) % 4).astype(np.uint32), sc.matrices, sc.should_callback, sc.entities)
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, scene, trees)
    assert st["n_hits"] > 500
    # (a) port on the exported refitted trees: bit-exact
    p_trees = [port.tree_import(t.export()) for t in trees]
    pres = oracle_frame(port, scene, p_trees, port=port)
    compare_frame(pres, st, bp, ep, hits, rel_of=lambda k: port.pair_matrix(scene.matrices[k[0]], scene.matrices[k[1]]))
    # (b) the reference rebuilding its trees from the re-posed triangles: tree-independent outputs
    o_trees = [oracle.tree_build(m.positions, m.normals, m.vertex_ids) for m in variants]
    ores = oracle_frame(oracle, scene, o_trees, port=port)

    def canon(pair, ta, tb):
        a, b = pair
        return (a, b, ta, tb) if a < b else (b, a, tb, ta)
    o_coll = {tuple(sorted(k)) for k, r in ores["per_pair"].items() if r.colliding}
    g_coll = {tuple(sorted((int(p["entry_first"]), int(p["entry_second"])))) for p in ep}
    o_hits = {canon(k, int(a), int(b)) for k, r in ores["per_pair"].items() for a, b in r.hit_ids.tolist()}
    g_hits = {canon(tuple(bp[h["pair"]].tolist()), int(h["tri_first"]), int(h["tri_second"])) for h in hits}
    r = dict(st=st, ref=ores["totals"], only_ref=o_hits - g_hits, only_gpu=g_hits - o_hits, coll_equal=(o_coll == g_coll),
             only_ref_coll=o_coll - g_coll, only_gpu_coll=g_coll - o_coll, n_hits=len(o_hits))
    _assert_tree_independent_parity(r, scene, port)
