"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (it reads /root/reference; the fixtures it writes travel):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference/src/emirdrp:
    processing/wavecal/apply_rectwv_coeff.py   (orchestration, paste, JMN/JMX scan, header)
    processing/wavecal/slitlet2d.py            (Slitlet2D: unpacking, crop, rectify call)
    processing/wavecal/set_wv_parameters.py    (grism/filter table)
    processing/wavecal/median_slitlets_rectified.py  (modes 0, 1: numpy only; mode 2 uses
                                                numina's define_mask_borders -> restated)
    core/__init__.py                           (EMIR constants)

What cannot run here and is replaced by shims (numina and astropy are not installed and
their sources are not under /root/reference):
    numina.array.wavecalib.{apply_integer_offsets, fix_pix_borders, resample},
    numina.array.distortion                -> oracle/numina_restated.py (restatement)
    astropy.io.fits, numina.frame.utils    -> pyemir_b200.hdu (Header / HDUList / copy_img)
    emirdrp.instrument.components.dtu      -> pyemir_b200.dtu  (pinned separately against the
                                              reference's test_dtu.py known answers)
    emirdrp.products.RectWaveCoeff         -> pyemir_b200.products

So these vectors pin everything pyemir itself does on the path (SURVEY.md §8 a1, a3, a4, a7,
a8, a11) against the reference's own code; the arithmetic delegated to numina (a2, a5, a6)
stays a restatement ("parity unpinned" for those rows).
"""

import hashlib
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/emirdrp"
sys.path.insert(0, ROOT)

from oracle import numina_restated as nr           # noqa: E402
from pyemir_b200 import dtu, hdu, products, synthetic  # noqa: E402


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _package(name, path=None):
    m = _module(name)
    m.__path__ = [path] if path else []
    return m


def install_shims():
    """Make `import emirdrp.processing.wavecal.apply_rectwv_coeff` work without numina/astropy."""
    # third-party pieces that are absent
    _package("astropy")
    _package("astropy.io")
    _module("astropy.io.fits", Header=hdu.Header, PrimaryHDU=hdu.PrimaryHDU,
            ImageHDU=hdu.ImageHDU, HDUList=hdu.HDUList)
    sys.modules["astropy.io"].fits = sys.modules["astropy.io.fits"]
    for pkg in ("numina", "numina.array", "numina.array.display", "numina.array.wavecalib",
                "numina.frame", "numina.tools"):
        _package(pkg)
    _module("numina.array.display.logging_from_debugplot",
            logging_from_debugplot=lambda debugplot: None)
    _module("numina.array.display.pause_debugplot",
            DEBUGPLOT_CODES=(0, -1, 1, -2, 2, -10, 10, -11, 11, -12, 12, -21, 21, -22, 22),
            pause_debugplot=lambda *a, **k: None)
    _module("numina.array.display.ximshow", ximshow=lambda *a, **k: None)
    _module("numina.array.wavecalib.apply_integer_offsets",
            apply_integer_offsets=nr.apply_integer_offsets)
    _module("numina.array.wavecalib.fix_pix_borders", find_pix_borders=nr.find_pix_borders,
            define_mask_borders=nr.define_mask_borders)
    _module("numina.array.wavecalib.resample", resample_image2d_flux=nr.resample_image2d_flux)
    _module("numina.array.distortion", order_fmap=nr.order_fmap, rectify2d=nr.rectify2d,
            fmap=nr.fmap)
    _module("numina.frame.utils", copy_img=hdu.copy_img)
    _module("numina.tools.arg_file_is_new", arg_file_is_new=lambda parser, arg, mode="wt": arg)
    # the reference package: real directories, but none of the __init__ files that pull in
    # numina's recipe machinery
    _package("emirdrp", REF)
    _package("emirdrp.processing", os.path.join(REF, "processing"))
    _package("emirdrp.processing.wavecal", os.path.join(REF, "processing", "wavecal"))
    _package("emirdrp.instrument")
    _package("emirdrp.instrument.components")
    _module("emirdrp.instrument.components.dtu", DtuConf=dtu.DtuConf)
    _module("emirdrp.products", RectWaveCoeff=products.RectWaveCoeff)
    _module("emirdrp.datamodel")                      # only used by the CLI entry points
    _module("emirdrp.instrument.csu_configuration", CsuConfiguration=object)
    # emirdrp.core is a plain constants module: load the real one
    spec = importlib.util.spec_from_file_location("emirdrp.core",
                                                  os.path.join(REF, "core", "__init__.py"))
    core = importlib.util.module_from_spec(spec)
    sys.modules["emirdrp.core"] = core
    spec.loader.exec_module(core)
    return importlib.import_module("emirdrp.processing.wavecal.apply_rectwv_coeff")


# ------------------------------------------------------------------ the cases
# (name, BASELINE config, seed offset, resampling, offsets, missing slitlets)
CASES = [
    ("j_area", 1, 0, 2, (0, 0), []),
    ("j_nearest", 1, 0, 1, (0, 0), []),
    ("j_offsets_missing", 1, 1, 2, (3, -2), [1, 27, 55]),
    ("h_area", 2, 0, 2, (0, 0), []),
    ("lr_yj_area", 4, 0, 2, (0, 0), [2]),
]
ROW_STEP, COL_STEP = 13, 41          # subsample kept in the fixture


def build_case(name, config, seed_off, resampling, offsets, missing):
    """Seeded input of one case: (HDUList, RectWaveCoeff).  Shared with the tests."""
    coeff = synthetic.make_config(config)
    coeff.global_integer_offset_x_pix = int(offsets[0])
    coeff.global_integer_offset_y_pix = int(offsets[1])
    coeff.missing_slitlets = list(missing)
    seed = synthetic.BASELINE_CONFIGS[config]["seed"] + seed_off
    scene = synthetic.make_base_scene(coeff, seed=seed)
    frame = synthetic.make_frames(coeff, 1, seed=seed, scene=scene)[0]
    hdul = synthetic.make_hdulist(frame.astype(np.float32), coeff)
    return hdul, coeff


def cards_of(header):
    """Every card except the time-stamped HISTORY line, as [keyword, value, comment]."""
    out = []
    for keyword, value, comment in header.cards:
        if keyword == "HISTORY" and str(value).startswith(
                "Rectification and wavelength calibration time"):
            continue
        if isinstance(value, (np.integer,)):
            value = int(value)
        elif isinstance(value, (np.floating,)):
            value = float(value)
        out.append([keyword, value, comment])
    return out


def median_input(ref):
    """Input of the median_slitlets_rectified vectors: the reference's rectified image of
    case j_area with a NaN, exact ties and an all-zero slitlet planted in it."""
    hdul, coeff = build_case(*CASES[0])
    out = ref.apply_rectwv_coeff(hdul, coeff, args_resampling=2)
    data = out[0].data
    data[38 * 4 + 7, 1500] = np.nan                       # slitlet 5: one NaN column
    data[38 * 9:38 * 9 + 38, 1200:1300] = np.float32(3.25)   # slitlet 10: ties
    data[38 * 19:38 * 20, :] = 0.0                        # slitlet 20: empty
    data[38 * 29:38 * 29 + 19, 1000] = -0.0               # signed zeros
    return out


def abba_vectors():
    import ast
    path = os.path.join(REF, "recipes", "spec", "abba.py")
    tree = ast.parse(open(path).read())
    func = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_isky")
    namespace = {}
    exec(compile(ast.Module(body=[func], type_ignores=[]), path, "exec"), namespace)
    get_isky = namespace["get_isky"]
    out = {}
    for basic_pattern in ("AB", "ABBA"):
        for repeat in (1, 2, 3):
            nimages = len(basic_pattern) * repeat * 3
            out[f"{basic_pattern}_{repeat}"] = [get_isky(i, basic_pattern, repeat)
                                                for i in range(nimages)]
    try:
        get_isky(0, "ABAB", 1)
    except ValueError as exc:
        out["error"] = str(exc)
    return out


def plain(v):
    """JSON-friendly copy of one set_wv_parameters value."""
    if hasattr(v, "coef"):                                   # numpy Polynomial
        return {"Polynomial": [float(c) for c in v.coef]}
    if isinstance(v, (np.ndarray, list, tuple)):
        return np.asarray(v).tolist()
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return float(v)
    return v


def main():
    ref = install_shims()
    swp = importlib.import_module("emirdrp.processing.wavecal.set_wv_parameters")
    core = sys.modules["emirdrp.core"]

    # ---- set_wv_parameters: every grism x filter combination, value or error
    table = {}
    for grism in list(core.EMIR_VALID_GRISMS) + ["X"]:
        for filt in list(core.EMIR_VALID_FILTERS) + ["X"]:
            try:
                wv = swp.set_wv_parameters(filt, grism)
                table[f"{grism}+{filt}"] = {k: plain(v) for k, v in sorted(wv.items())}
            except ValueError as exc:
                table[f"{grism}+{filt}"] = {"ValueError": str(exc)}
    constants = {k: getattr(core, k) for k in
                 ("EMIR_NBARS", "EMIR_NAXIS1", "EMIR_NAXIS2", "EMIR_NAXIS1_ENLARGED",
                  "EMIR_NPIXPERSLIT_RECTIFIED", "EMIR_VALID_GRISMS", "EMIR_VALID_FILTERS")}
    with open(os.path.join(HERE, "set_wv_parameters.json"), "w") as fh:
        json.dump({"source": "emirdrp/processing/wavecal/set_wv_parameters.py + core/__init__.py "
                             "(reference code, executed)",
                   "constants": constants, "table": table}, fh, indent=1, sort_keys=True)

    # ---- apply_rectwv_coeff: reference orchestration on seeded inputs
    meta = {"source": "emirdrp/processing/wavecal/apply_rectwv_coeff.py + slitlet2d.py "
                      "(reference code, executed; numina numerics from oracle/numina_restated.py)",
            "row_step": ROW_STEP, "col_step": COL_STEP, "cases": {}}
    arrays = {}
    for name, config, seed_off, resampling, offsets, missing in CASES:
        hdul, coeff = build_case(name, config, seed_off, resampling, offsets, missing)
        before = hdul[0].data.copy()
        out = ref.apply_rectwv_coeff(hdul, coeff, args_resampling=resampling)
        assert np.array_equal(hdul[0].data, before), "the reference must not mutate its input"
        data = out[0].data
        assert data.dtype == np.float32
        meta["cases"][name] = {
            "config": config, "seed_offset": seed_off, "resampling": resampling,
            "offsets": list(offsets), "missing": list(missing),
            "shape": list(data.shape), "dtype": str(data.dtype),
            "sha256": hashlib.sha256(np.ascontiguousarray(data).tobytes()).hexdigest(),
            "sum": float(data.astype(np.float64).sum()),
            "cards": cards_of(out[0].header),
        }
        arrays[name] = np.ascontiguousarray(data[::ROW_STEP, ::COL_STEP])
        print(name, data.shape, meta["cases"][name]["sha256"][:16], arrays[name].shape)

    # ---- error behaviour of the reference orchestration
    errors = {}
    hdul, coeff = build_case(*CASES[0])
    hdul[0].header["filter"] = "H"
    try:
        ref.apply_rectwv_coeff(hdul, coeff)
    except ValueError as exc:
        errors["filter_mismatch"] = str(exc)
    hdul, coeff = build_case(*CASES[0])
    hdul[0].header["grism"] = "K"
    try:
        ref.apply_rectwv_coeff(hdul, coeff)
    except ValueError as exc:
        errors["grism_mismatch"] = str(exc)
    hdul, coeff = build_case(*CASES[0])
    hdul[0].header["xdtu"] = float(hdul[0].header["xdtu"]) + 0.5
    try:
        ref.apply_rectwv_coeff(hdul, coeff, args_ignore_dtu_configuration=False)
    except ValueError as exc:
        errors["dtu_mismatch"] = str(exc)
    hdul, coeff = build_case(*CASES[0])
    try:
        ref.apply_rectwv_coeff(hdul, coeff, args_resampling=3)
    except ValueError as exc:
        errors["bad_resampling"] = str(exc)
    meta["errors"] = errors
    print(errors)

    # ---- median_slitlets_rectified: the reference function on the rectified j_area image,
    #      with a NaN, ties and an all-zero slitlet planted in it
    med = importlib.import_module("emirdrp.processing.wavecal.median_slitlets_rectified")
    med_in = median_input(ref)
    meta["median"] = {}
    for mode, useful in ((0, None), (1, None), (2, None), (2, list(range(3, 50)))):
        out = med.median_slitlets_rectified(med_in, mode=mode, list_useful_slitlets=useful)
        data = out[0].data
        assert data.dtype == np.float32
        key = f"mode{mode}" + ("" if useful is None else "_useful3to49")
        meta["median"][key] = {
            "mode": mode, "useful": useful, "shape": list(data.shape),
            "sha256": hashlib.sha256(np.ascontiguousarray(data).tobytes()).hexdigest()}
        arrays["median_" + key] = (np.ascontiguousarray(data[::ROW_STEP, ::COL_STEP])
                                   if data.ndim == 2 and data.shape[0] > 55 else data[..., ::7])
        print("median", key, data.shape, meta["median"][key]["sha256"][:16])
    for what, bad in (("naxis2", lambda h: setattr(h[0], "data", h[0].data[:100])),
                      ("instrume", lambda h: h[0].header.__setitem__("instrume", "OSIRIS"))):
        hdul = hdu.copy_img(med_in)
        bad(hdul)
        try:
            med.median_slitlets_rectified(hdul)
        except ValueError as exc:
            meta["median"]["error_" + what] = str(exc)

    # ---- ABBA bookkeeping: the reference's own get_isky (recipes/spec/abba.py:46-75), taken
    #      out of the module source because the module itself needs numina / astropy / scipy
    meta["abba"] = abba_vectors()

    with open(os.path.join(HERE, "apply_rectwv_coeff.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "apply_rectwv_coeff_subsample.npz"), **arrays)


if __name__ == "__main__":
    main()
