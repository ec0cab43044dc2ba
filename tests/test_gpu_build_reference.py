"""GPU, IMRCD_BUILD_REFERENCE: the device build must reproduce the reference's tree bit for bit (rows B1-B5 of SURVEY 8a:
OBB fit with sequential FP64 sums + eig3 + rows-of-V axes, split rule, child order, leaf/triangle order)."""
import numpy as np
import pytest

import golden_io
from inmyroom_vulkan_b200 import scenes
from inmyroom_vulkan_b200.collision import IMRCD_BUILD_REFERENCE, CollisionDetection, OBBtree
from helpers import compare_frame, f32_bits, gpu_frame, oracle_frame

pytestmark = pytest.mark.gpu


def assert_same_tree(flat, gold):
    for f in golden_io.TREE_FIELDS:
        g = np.asarray(getattr(gold, f)); o = np.asarray(getattr(flat, f))
        assert o.shape == g.shape, f"{f}: shape {o.shape} vs {g.shape}"
        if g.dtype == np.float32:
            assert np.array_equal(f32_bits(o), f32_bits(g)), f"{f} differs in bits"
        else:
            assert np.array_equal(o, g), f


def test_obb_fit_golden(gpu_ctx):
    z = golden_io.load("obb_fit")
    off = 0
    for k, cnt in enumerate(z["sizes"].tolist()):
        box = gpu_ctx.test_obb_fit(z["points"][off:off + cnt]); off += cnt
        assert np.array_equal(f32_bits(box), f32_bits(z["boxes"][k])), f"cloud {k}"


def test_obb_fit_random(gpu_ctx, oracle):
    rng = np.random.default_rng(31)
    for k in range(120):
        cnt = int(rng.integers(1, 400))
        p = (rng.normal(size=(cnt, 3)) * (rng.random(3) * 100 + 0.01) + rng.normal(size=3) * 1000).astype(np.float32)
        assert np.array_equal(f32_bits(gpu_ctx.test_obb_fit(p)), f32_bits(oracle.obb_from_points(p))), k


@pytest.mark.parametrize("name", golden_io.tree_names())
def test_reference_tree_golden(gpu_ctx, name):
    gold, mesh = golden_io.golden_tree(golden_io.load("trees"), name)
    flat = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids, build_mode=IMRCD_BUILD_REFERENCE).export()
    assert_same_tree(flat, gold)


@pytest.mark.parametrize("mesh", [scenes.torus(100, 50), scenes.uv_sphere(66, 65), scenes.grid_sheet(60, 40, 1500.0, 900.0, bump=30.0),
                                  scenes.box_mesh(1, 2, 3, sub=8), scenes.cylinder(24, 12, 27.5, 137.5)], ids=lambda m: m.name)
def test_reference_tree_identical_to_oracle(gpu_ctx, oracle, mesh):
    flat = OBBtree(gpu_ctx, mesh.positions, mesh.normals, mesh.vertex_ids, build_mode=IMRCD_BUILD_REFERENCE).export()
    assert_same_tree(flat, oracle.tree_build(mesh.positions, mesh.normals, mesh.vertex_ids).flat)


def test_reference_tree_without_normals_and_ids(gpu_ctx, oracle):
    mesh = scenes.torus(30, 12)
    flat = OBBtree(gpu_ctx, mesh.positions, None, None, build_mode=IMRCD_BUILD_REFERENCE).export()
    assert_same_tree(flat, oracle.tree_build(mesh.positions, None, None).flat)


def test_frame_with_reference_built_trees(gpu_ctx, oracle):
    """End to end with no imported data at all: GPU-built reference trees + GPU frame == the reference, bit for bit."""
    static = scenes.atrium_static(detail=1)
    keep = [0, 3, 8, 9, 64, 65, 120, 125]
    static = ([static[0][i] for i in keep], static[1][keep])
    scene = scenes.scene_static_vs_bodies(scenes.uv_sphere(24, 17), 200, seed=9, body_scale=(0.5, 1.5), static=static)
    g_trees = [OBBtree(gpu_ctx, m.positions, m.normals, m.vertex_ids, build_mode=IMRCD_BUILD_REFERENCE) for m in scene.meshes]
    o_trees = [oracle.tree_build(m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    cd = CollisionDetection(ctx=gpu_ctx)
    st, bp, ep, hits = gpu_frame(cd, scene, g_trees)
    ores = oracle_frame(oracle, scene, o_trees)
    compare_frame(ores, st, bp, ep, hits, rel_of=lambda k: oracle.pair_matrix(scene.matrices[k[0]], scene.matrices[k[1]]))
    assert st["n_hits"] > 0
