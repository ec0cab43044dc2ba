"""SURVEY 8f F3, the game itself: SnakeGame's own SnakePlayerComp under the engine's own frame loop, reference against drop-in.

oracle/_ref/libsnake_game.so is the game's REAL DLL (testGames/SnakeGame/game_dll and the engine sources its CMakeLists lists, unmodified,
-DGAME_DLL); oracle/_ref/snake_harness_{ref,dropin} are ONE headless engine (tests/cpp/snake_harness.cpp) around the reference's real
ECSwrapper and general components (NodeData, AnimationComposer, AnimationActor, Early/LateNodeGlobalMatrix, Camera, ModelCollision), once
with the reference's CollisionDetection and once with the drop-in over libimrcd.so (ModelCollisionComp.cpp compiles unchanged against it).
The loop is closed: ModelCollisionComp::Update makes the entries from the node matrices, the deltaVectors come back through
SnakePlayerComp::CollisionCallback and move the snakes (SnakePlayerCompEntity.cpp:223-251), the moved snakes make the next frame's entries.
Both runs see the same fixed 1/60 s clock, so they stay on the same trajectory as long as the deltaVectors agree."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
REF = os.path.join(REFDIR, "snake_harness_ref")
DROP = os.path.join(REFDIR, "snake_harness_dropin")
SHADOW = os.path.join(REFDIR, "snake_harness_shadow")
GAME = os.path.join(REFDIR, "libsnake_game.so")


def _run(exe, frames, snakes):
    r = subprocess.run([exe, str(frames), str(snakes), GAME], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    callbacks, pos = [], []
    for ln in r.stdout.splitlines():
        w = ln.split()
        if w[0] == "frame":
            callbacks.append(int(w[3])); pos.append([])
        elif w[0] == "snake":
            pos[-1].append([float(x) for x in w[2:5]])
    return np.array(callbacks), np.array(pos)


def _build_if_possible(*targets):
    if os.path.isdir("/root/reference/inMyRoom_vulkan"):
        subprocess.run(["make", "-s", "-j4", "-C", os.path.join(ROOT, "oracle"), *targets], check=True)


def test_reference_game_runs_headless():
    """CPU twin alone: the real game DLL under the reference's own collision detection -- snakes fall, land on the floor (the response
    stage holds them there), wander and bump into things; two runs are identical (fixed clock, std::rand unseeded)."""
    _build_if_possible(GAME, REF)
    if not (os.path.exists(REF) and os.path.exists(GAME)):
        pytest.skip("oracle/_ref/snake_harness_ref not available")
    cb, pos = _run(REF, 150, 4)
    cb2, pos2 = _run(REF, 150, 4)
    assert np.array_equal(cb, cb2) and np.array_equal(pos, pos2)
    assert cb.sum() > 100 and pos.shape == (150, 4, 3)
    # +y is down: gravity alone would have taken a snake ~ 0.5 * 9.8 * 2.5^2 = 30 units below its start; the floor's top face is y = 0
    assert pos[-1, :, 1].max() < 0.5 and np.abs(pos[-1] - pos[0]).max() > 0.5


def _shadow(frames, snakes):
    r = subprocess.run([SHADOW, str(frames), str(snakes), GAME], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("shadow")]
    assert len(rows) == frames
    # shadow <frame> compared <n> mismatched <n> worst <rel> abs <abs> over <n>
    return dict(compared=int(rows[-1][3]), mismatched=int(rows[-1][5]), worst_rel=max(float(w[7]) for w in rows),
                worst_abs=max(float(w[9]) for w in rows), over=int(rows[-1][11]))


@pytest.mark.gpu
@pytest.mark.parametrize("frames,snakes", [(300, 6), (200, 12)])
def test_snake_game_lock_step(gpu_ctx, frames, snakes):
    """The rigorous one.  snake_harness_shadow puts BOTH implementations behind the one class the engine's ModelCollisionComp talks to: the
    reference's callbacks drive the game, the drop-in receives the very same entries every frame and its callbacks are compared with the
    reference's in lock step -- same receivers (ancestor rule), same (familyEntity, collideWithEntity), deltaVectors within 1e-4 of their
    length + 2e-5 world units (the game's resting contacts produce deltaVectors of ~1e-3 units from ray hits at coordinates of ~10 units:
    a few FP32 ulps of the coordinates is all the precision such a delta has, whatever the order of the sums behind the ray origins)."""
    _build_if_possible("harness")
    if not (os.path.exists(SHADOW) and os.path.exists(GAME)):
        pytest.skip("oracle/_ref/snake_harness_shadow not available")
    s = _shadow(frames, snakes)
    assert s["compared"] > 10000 and s["mismatched"] == 0 and s["over"] == 0, s
    assert s["worst_abs"] < 2e-5, s


def test_free_running_note():
    """Why the free-running comparison below is loose: the closed loop is chaotic.  A deltaVector that differs in its last bits moves a
    snake by 1e-8 units, a triangle pair that was just touching stops touching, and from there the two games are different games (two
    runs of the drop-in differ from each other the same way).  test_snake_game_lock_step is the parity statement."""


@pytest.mark.gpu
@pytest.mark.parametrize("frames,snakes", [(120, 6)])
def test_snake_game_on_the_drop_in_follows_the_reference(gpu_ctx, frames, snakes):
    _build_if_possible("harness")
    if not (os.path.exists(REF) and os.path.exists(DROP) and os.path.exists(GAME)):
        pytest.skip("oracle/_ref/snake_harness_* not available")
    cb_w, pos_w = _run(REF, frames, snakes)
    cb_g, pos_g = _run(DROP, frames, snakes)
    assert pos_g.shape == pos_w.shape == (frames, snakes, 3) and cb_w.sum() > 100
    err = np.abs(pos_g - pos_w).max(axis=(1, 2))
    # the first frames (before any contact can flip, see test_free_running_note) are the same game to FP32 rounding ...
    assert err[:20].max() < 1e-4, (int(err[:20].argmax()), float(err[:20].max()))
    assert np.array_equal(cb_g[:20], cb_w[:20])
    # ... and afterwards it is at least the same kind of game: every snake the reference keeps on the floor (top face y = 0, +y is down)
    # is kept there by the drop-in too, and none has been pushed through it
    on_floor_w = pos_w[-1, :, 1] > -1.0
    assert np.array_equal(pos_g[-1, :, 1] > -1.0, on_floor_w) and pos_g[-1, on_floor_w, 1].max() < 0.5
