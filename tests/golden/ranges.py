"""Synthetic-input value ranges per solution (var-name prefix -> (lo, hi)) and the table of emitter-generated solutions
with their reference build tags.  Pure data: imported by the fixture generator, the GPU tests and bench_stencils.py
(nothing here touches oracle/)."""

# (var -> (lo, hi)) value ranges per stencil; SURVEY.md section 8(d).
# Keys are var-name PREFIXES (longest match wins).  awp/ssg ranges are chosen so that a step changes the
# fields by O(10 %) (the reference's own init_diff inputs blow up, SURVEY.md section 7 hard part 3).
RANGES = {
    "iso3dfd": {"p": (-1.0, 1.0), "v": (0.05, 0.3)},
    "awp_elastic": {"vel": (-1.0, 1.0), "stress": (-1.0, 1.0), "rho": (1.0, 2.0), "mu": (1.0, 3.0), "lambda": (1.0, 3.0),
                    "delta_t": (1e-2, 1e-2), "h": (1.0, 1.0), "cr": (0.92, 1.0)},
    "ssg": {"v_": (-1e-3, 1e-3), "s_": (-1e-3, 1e-3), "rho": (1.0, 2.0), "mu": (1.0, 3.0), "lambda": (1.0, 3.0), "lambdamu2": (1.0, 3.0)},
    "awp": {"vel": (-1.0, 1.0), "stress": (-1.0, 1.0), "rho": (1.0, 2.0), "mu": (1.0, 3.0), "lambda": (1.0, 3.0),
            "delta_t": (1e-2, 1e-2), "h": (1.0, 1.0), "cr": (0.92, 1.0), "weight": (0.1, 1.0), "tau2": (0.1, 0.9), "anelastic": (0.1, 1.0)},
    "iso3dfd_sponge": {"p": (-1.0, 1.0), "v": (0.05, 0.3), "cr": (0.92, 1.0)},
    "awp_elastic_abc": {"vel": (-1.0, 1.0), "stress": (-1.0, 1.0), "rho": (1.0, 2.0), "mu": (1.0, 3.0), "lambda": (1.0, 3.0),
                        "delta_t": (1e-2, 1e-2), "h": (1.0, 1.0), "cr": (0.92, 1.0)},
    "awp_abc": {"vel": (-1.0, 1.0), "stress": (-1.0, 1.0), "rho": (1.0, 2.0), "mu": (1.0, 3.0), "lambda": (1.0, 3.0),
                "delta_t": (1e-2, 1e-2), "h": (1.0, 1.0), "cr": (0.92, 1.0), "weight": (0.1, 1.0), "tau2": (0.1, 0.9), "anelastic": (0.1, 1.0)},
    "tti": {"u": (-1.0, 1.0), "v": (-1.0, 1.0), "m": (1.0, 2.0), "damp": (0.0, 0.1), "phi": (0.0, 1.0), "theta": (0.0, 1.0),
            "delta": (0.1, 0.3), "epsilon": (0.1, 0.3), "ti": (0.1, 1.0)},
    "3axis": {"A": (-1.0, 1.0)}, "3axis_with_diags": {"A": (-1.0, 1.0)}, "3plane": {"A": (-1.0, 1.0)}, "cube": {"A": (-1.0, 1.0)},
}
# solutions produced by the CUDA emitter (yask_b200/csrc/gen/manifest.json) and their reference build tags
GENERATED = {"awp_elastic": "awp_elastic", "ssg": "ssg-fp64", "awp": "awp", "iso3dfd_sponge": "iso3dfd_sponge", "tti": "tti",
             "3axis": "3axis", "3axis_with_diags": "3axis_with_diags", "3plane": "3plane", "cube": "cube",
             "awp_elastic_abc": "awp_elastic_abc", "awp_abc": "awp_abc", "test_1d": "test_1d", "test_2d": "test_2d", "test_3d": "test_3d",
             "test_boundary_3d": "test_boundary_3d", "test_stream_3d": "test_stream_3d"}
GENERATED["iso3dfd_fp64"] = "iso3dfd-fp64"
RANGES["iso3dfd_fp64"] = RANGES["iso3dfd"]
# the reference's own validation matrix for these two (src/kernel/Makefile:1155-1156): fp64, radius 3 / radius 6
GENERATED["iso3dfd_fp64_r3"] = "iso3dfd-fp64-r3"
RANGES["iso3dfd_fp64_r3"] = RANGES["iso3dfd"]
GENERATED["iso3dfd_sponge_fp64_r6"] = "iso3dfd_sponge-fp64-r6"
RANGES["iso3dfd_sponge_fp64_r6"] = RANGES["iso3dfd_sponge"]
GENERATED.update({"fsg": "fsg", "fsg_abc": "fsg_abc", "ssg2": "ssg2", "ssg_merged": "ssg_merged", "fsg2": "fsg2"})
for _t in ("ssg2", "ssg_merged", "fsg2"):
    # merged-array variants: v/s hold all components along a misc dim, coef the material coefficients
    RANGES[_t] = {"v": (-1e-3, 1e-3), "s": (-1e-3, 1e-3), "rho": (1.0, 2.0), "coef": (1.0, 3.0), "c": (1.0, 3.0), "": (0.5, 1.0)}
for _t in ("fsg", "fsg_abc"):
    # FSG elastic: velocities/stresses small, material coefficients (c11..c66) O(1), density O(1)
    RANGES[_t] = {"v_": (-1e-3, 1e-3), "s_": (-1e-3, 1e-3), "rho": (1.0, 2.0), "c": (1.0, 3.0), "": (0.5, 1.0)}
for _t in ("test_1d", "test_2d", "test_3d", "test_boundary_3d", "test_stream_3d"):
    RANGES[_t] = {"": (-1.0, 1.0)}       # every var of the reference's test solutions: [-1, 1)


# second batch: filters, FSG variants, 1-D/2-D tests, stages, scratch vars, step conditions,
# math functions
_TESTS2 = ("test_boundary_1d", "test_boundary_2d", "test_partial_3d", "test_stages_1d", "test_stages_2d", "test_stages_3d",
           "test_stream_1d", "test_stream_2d", "test_reverse_2d", "test_scratch_1d", "test_scratch_2d", "test_scratch_3d",
           "test_scratch_boundary_1d", "test_scratch_stages_1d", "test_step_cond_1d", "test_func_1d", "box_filter", "gaussian_filter")
for _t in _TESTS2:
    GENERATED[_t] = _t
    RANGES[_t] = {"": (-1.0, 1.0)}
RANGES["test_func_1d"] = {"": (0.5, 1.5)}        # cbrt/atan arguments away from zero
for _t in ("fsg2_abc", "fsg_merged", "fsg_merged_abc"):
    GENERATED[_t] = _t
    RANGES[_t] = {"v": (-1e-3, 1e-3), "s": (-1e-3, 1e-3), "rho": (1.0, 2.0), "coef": (1.0, 3.0), "c": (1.0, 3.0), "": (0.5, 1.0),
                  "sponge": (0.9, 1.0)}

# 2-D wave / shallow-water solutions (scratch-var chains): bounded amplitudes, positive depth, unit grid
GENERATED["wave2d"] = "wave2d"
RANGES["wave2d"] = {"u": (-0.1, 0.1), "v": (-0.1, 0.1), "e": (-0.1, 0.1), "dt": (0.01, 0.01), "inv_d": (1.0, 1.0), "g": (9.81, 9.81),
                    "depth": (1.0, 1.0)}
GENERATED["swe2d"] = "swe2d"
RANGES["swe2d"] = {"u": (-0.1, 0.1), "v": (-0.1, 0.1), "e": (-0.1, 0.1), "q": (-0.1, 0.1), "pe": (-0.1, 0.1), "keH": (-0.1, 0.1),
                   "h": (1.0, 2.0), "dt": (1e-3, 1e-3), "dx": (1.0, 1.0), "dy": (1.0, 1.0), "inv_d": (1.0, 1.0), "g": (9.81, 9.81),
                   "coriolis": (0.1, 0.1), "pe_offset": (0.0, 0.0), "ti_exp": (1.0, 1.0)}


def range_of(ranges, name):
    best = None
    for k, v in ranges.items():
        if name.startswith(k) and (best is None or len(k) > len(best[0])):
            best = (k, v)
    return best[1]
