#!/usr/bin/env python
"""Generates tests/golden/*.json: seeded inputs (by generator + seed, not stored) and the outputs
of the CPU oracle on them, stored bit-exactly (doubles as hex strings).

  python tests/golden/make_golden.py

The vectors are written by the ORACLE: tests/test_golden.py checks on the CPU that the oracle
still reproduces them and, on the GPU, that the CUDA path does — a drift of either shows up
against a fixed, committed answer instead of only against each other.  The RANSAC, residual and
P6L entries are ALSO what the reference's own sources give (compiled against Eigen / glog
stand-ins by oracle/build_ref.sh; tests/test_ref_p6l.py::
test_reference_reproduces_the_golden_vectors); config2_reference.json is written by that
reference build directly (make_config2_reference.py).  The BA entries stay oracle-only (Ceres'
solver cannot be built here)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle as O                                              # noqa: E402
from privacy_preserving_sfm_b200 import synthetic as S         # noqa: E402

RANSAC_CASES = [
    # name, scene kwargs, options (max_error, min_inlier_ratio, confidence, multiplier, min, max)
    ("mapper_settings", dict(n=2000, inlier_ratio=0.3, seed=101),
     (0.012, 0.25, 0.99999, 3.0, 100, 10000)),
    ("fixed_3000_trials", dict(n=4000, inlier_ratio=0.35, seed=102),
     (0.012, 0.25, 0.99999, 3.0, 3000, 3000)),
    ("early_abort", dict(n=1500, inlier_ratio=0.7, seed=103),
     (0.012, 0.1, 0.99, 3.0, 0, 2 ** 64 - 1)),
    ("several_waves_then_abort", dict(n=3000, inlier_ratio=0.45, seed=104),
     (0.012, 0.25, 0.99999, 3.0, 2048, 10000)),
]


def hexes(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def ransac_case(scene_kw, opt):
    sc = S.make_abs_pose_scene(**scene_kw)
    O.set_prng_seed(0)
    rep, mask = O.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], O.make_options(*opt))
    return {
        "success": int(rep.success), "num_trials": int(rep.num_trials),
        "num_inliers": int(rep.num_inliers), "residual_sum": float(rep.residual_sum).hex(),
        "best_trial": int(rep.best_trial), "best_model_idx": int(rep.best_model_idx),
        "num_models_scored": int(rep.num_models_scored), "model": hexes(list(rep.model)),
        "mask_sha256": hashlib.sha256(np.asarray(mask, np.uint8).tobytes()).hexdigest(),
        "prng_peek_after": int(O.prng_peek()),
    }


def residual_case():
    sc = S.make_abs_pose_scene(n=500, inlier_ratio=0.5, seed=105)
    P = np.concatenate([sc["R"].T.reshape(9), sc["t"]])      # column-major 3x4 of the true pose
    r = O.line_residuals(sc["lines"], sc["points"], P)
    return {"residuals_sha256": hashlib.sha256(np.asarray(r, np.float64).tobytes()).hexdigest(),
            "first8": hexes(r[:8]), "num_le_thr": int((r <= 0.012 ** 2).sum())}


def p6l_case():
    probs = S.make_p6l_minimal_problems(8, seed=106)
    out = []
    for p in probs:
        sols = O.p6l_estimate(p["lines"], np.zeros(6, np.uint8), p["points"])
        out.append([hexes(s) for s in sols])
    return out


def ba_case():
    sb = S.make_ba_scene(num_cams=6, num_points=200, obs_per_point=4, seed=107)
    flags = np.zeros(6, np.uint8)
    flags[0], flags[1] = 1, 2          # camera 0 constant, camera 1 tvec[0] constant
    a = O.BaArrays(sb["qvecs"], sb["tvecs"], sb["points"], sb["obs_cam"], sb["obs_pt"],
                   sb["obs_line"], [1], [sb["cam_params"]], pose_flags=flags)
    ok, s = O.ba_solve(a, O.ba_default_options(num_threads=1, max_num_iterations=20,
                                               gradient_tolerance=1e-4))
    state = np.concatenate([a.qvecs.ravel(), a.tvecs.ravel(), a.points.ravel()])
    return {"ok": bool(ok), "successful_steps": int(s.num_successful_steps),
            "unsuccessful_steps": int(s.num_unsuccessful_steps),
            "termination_type": int(s.termination_type),
            "initial_cost": float(s.initial_cost).hex(), "final_cost": float(s.final_cost).hex(),
            "state_sha256": hashlib.sha256(state.astype(np.float64).tobytes()).hexdigest()}


def ba_intrinsics_case():
    """refine_extra_params on two SIMPLE_RADIAL cameras, the second in ConstantCameras()
    (ParameterizeCameras, src/optim/bundle_adjustment.cc:490-528); noisy scene, 5 LM iterations."""
    sb = S.make_ba_scene(num_cams=6, num_points=200, obs_per_point=4, seed=107, noise_px=2.0)
    flags = np.zeros(6, np.uint8)
    flags[0], flags[1] = 1, 2
    prm = [[1000.0, 500.0, 500.0, 0.08], [1000.0, 500.0, 500.0, 0.05]]
    a = O.BaArrays(sb["qvecs"], sb["tvecs"], sb["points"], sb["obs_cam"], sb["obs_pt"],
                   sb["obs_line"], [2, 2], prm, image_camera=np.arange(6) % 2, pose_flags=flags,
                   camera_const=[0, 1])
    ok, s = O.ba_solve(a, O.ba_default_options(num_threads=1, max_num_iterations=5,
                                               refine_extra_params=1))
    state = np.concatenate([a.qvecs.ravel(), a.tvecs.ravel(), a.points.ravel(),
                            a.camera_params.ravel()])
    return {"ok": bool(ok), "successful_steps": int(s.num_successful_steps),
            "unsuccessful_steps": int(s.num_unsuccessful_steps),
            "effective_parameters": int(s.num_effective_parameters_reduced),
            "initial_cost": float(s.initial_cost).hex(), "final_cost": float(s.final_cost).hex(),
            "k_camera0": float(a.camera_params[0, 3]).hex(),
            "k_camera1": float(a.camera_params[1, 3]).hex(),
            "state_sha256": hashlib.sha256(state.astype(np.float64).tobytes()).hexdigest()}


def main():
    O.build()
    gold = {"ransac": {name: dict(scene=kw, options=list(opt), expect=ransac_case(kw, opt))
                       for name, kw, opt in RANSAC_CASES},
            "line_residuals": residual_case(), "p6l_estimate": p6l_case(),
            "ba_solve_single_thread": ba_case(),
            "ba_intrinsics_single_thread": ba_intrinsics_case()}
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "oracle_vectors.json"))


if __name__ == "__main__":
    main()
