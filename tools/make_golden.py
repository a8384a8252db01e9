"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE in the build container (it cannot travel to the GPU box).

Usage:   eval $(tools/refenv/setup_ref.sh) && python tools/make_golden.py [--only NAME]

Two kinds of fixtures are written:
  * ref_goldens.npz  -- the numeric golden vectors the reference's own tests pin (tests/test_data/*.txt there),
                        re-saved verbatim with their provenance, so that the oracle can be checked against them where
                        /root/reference does not exist;
  * ref_*.npz        -- outputs of the reference itself, run here on the inputs named in each entry
                        (profiles, alpha_+, A^2, spectra, omgw0).
Nothing from the reference's sources is copied.
"""
import argparse
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_TEST_DATA = "/root/reference/tests/test_data"


def goldens():
    d = {}
    for name in ("shell", "shells_weak", "shells_inter", "shells_esp", "bubble", "data_compare_nuc-test_reference",
                 "xi-v_plane"):
        d[name.replace("-", "_")] = np.loadtxt(os.path.join(REF_TEST_DATA, name + ".txt"))
    for name in ("v_plus_minus", "v_minus_plus"):
        d[name] = np.loadtxt(os.path.join(REF_TEST_DATA, name + ".txt"))
    np.savez_compressed(os.path.join(OUT, "ref_goldens.npz"), **d)
    print("ref_goldens.npz", {k: v.shape for k, v in d.items()})


BAG_POINTS = [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.3, 0.2), (0.44, 0.0046), (0.7, 0.052), (0.58, 0.3),
              (0.92, 0.05), (0.731, 0.05), (0.05, 0.3), (0.2, 0.33), (0.6, 0.004), (0.56, 0.05), (0.8, 0.0046)]


def bag_shells():
    """sound_shell_bag profiles (decimated) + alpha_+ + scalars from the reference."""
    from pttools.bubble import fluid_bag, alpha, transition, quantities
    d = {"points": np.array(BAG_POINTS)}
    for i, (vw, an) in enumerate(BAG_POINTS):
        v, w, xi = fluid_bag.sound_shell_bag(vw, an)
        st = transition.identify_solution_type_bag(vw, an)
        ap = alpha.find_alpha_plus_bag(vw, an)
        d[f"v_{i}"], d[f"w_{i}"], d[f"xi_{i}"] = v, w, xi
        d[f"scal_{i}"] = np.array([ap, alpha.alpha_n_max_deflagration_bag(vw), quantities.ubarf_squared(v, w, xi, vw),
                                   {"Detonation": 2, "Hybrid": 1, "Subsonic deflagration": 0}[st.value]])
    np.savez_compressed(os.path.join(OUT, "ref_bag_shells.npz"), **d)
    print("ref_bag_shells.npz", len(BAG_POINTS))


SPEC_POINTS = [(0.5, 0.1), (0.7, 0.1), (0.9, 0.1), (0.44, 0.0046), (0.56, 0.05)]


def bag_spectra():
    """power_gw_scaled_bag / power_v_bag / a2_e_conserving_bag from the reference at reduced sizes."""
    import pttools.ssmtools as ssm
    z = np.logspace(-1, 3, 200)
    npt = (2000, 500, 1000)
    d = {"points": np.array(SPEC_POINTS), "z": z, "npt": np.array(npt)}
    zq = np.logspace(np.log10(0.2), np.log10(1000), 500)
    d["z_a2"] = zq
    for i, (vw, an) in enumerate(SPEC_POINTS):
        for nuc in ("exponential", "simultaneous"):
            d[f"pow_gw_{nuc[:3]}_{i}"] = ssm.power_gw_scaled_bag(z, (vw, an, ssm.NucType(nuc), (1.,)), npt=npt)
            d[f"pow_v_{nuc[:3]}_{i}"] = ssm.power_v_bag(z, (vw, an, ssm.NucType(nuc), (1.,)), npt=npt)
        a2, fp2, lam2 = ssm.a2_e_conserving_bag(zq, vw, an, npt=npt)
        d[f"a2_{i}"], d[f"fp2_{i}"], d[f"lam2_{i}"] = a2, fp2, lam2
    np.savez_compressed(os.path.join(OUT, "ref_bag_spectra.npz"), **d)
    print("ref_bag_spectra.npz", len(SPEC_POINTS))


def sin_transforms():
    """sin_transform on synthetic profiles (config 4 shape, tiny)."""
    import pttools.ssmtools as ssm
    rng = np.random.default_rng(0)
    z = np.logspace(-1, 3, 300)
    xi = np.linspace(0, 1, 513)
    d = {"z": z, "xi": xi}
    fs = []
    for p in range(4):
        phi, a, b = rng.uniform(0, 1), rng.uniform(0.1, 0.4), rng.uniform(0.5, 0.9)
        f = np.maximum(0, np.cos(xi + phi)) * ((xi > a) & (xi < b))
        fs.append(f)
        d[f"st_{p}"] = ssm.sin_transform(z, xi, f)
    d["f"] = np.array(fs)
    np.savez_compressed(os.path.join(OUT, "ref_sin_transform.npz"), **d)
    print("ref_sin_transform.npz")


TASKS = {"goldens": goldens, "bag_shells": bag_shells, "bag_spectra": bag_spectra, "sin_transforms": sin_transforms}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    try:
        import pttools  # noqa: F401
    except ImportError:
        sys.exit("the reference is not importable: eval $(tools/refenv/setup_ref.sh) first")
    for name, fn in TASKS.items():
        if args.only in (None, name):
            fn()
