"""Helpers shared by the CPU (oracle vs golden) and GPU (CUDA vs golden / oracle) parity tests."""
import datetime
import glob
import json
import os

import numpy as np

from sphy_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# golden state name -> attribute on the model objects (oracle port and SphyModel share the reference's names)
STATE_KEYS = ("RootWater", "SubWater", "CapRise", "RootDrain", "GwRecharge", "Gw", "BaseR", "H_gw", "SubDrain",
              "SnowStore", "SnowWatStore", "TotalSnowStore", "SoilWater", "QRAold", "StorRES", "RootRR", "RootDR",
              "RainR", "SnowR", "GlacR", "waterbalanceTot", "GlacFrac", "Scanopy")
COMP_KEYS = ("RootRRAold", "RootDRAold", "RainRAold", "SnowRAold", "GlacRAold", "BaseRAold")
# daily report file prefix -> key in the models' per-step `out` dict
REPORT_KEYS = {"TotR": "TotR", "ETr": "ETref", "ETaF": "ActETact", "Infil": "Infil", "wbal": "wbal",
               "DetRn": "DetRn", "DetRun": "DetRun", "SDpFld": "SDepFld", "STrans": "SedTrans", "Sdep": "SedDep", "SdYld": "SedYld",
               "SdFlux": "SedFlux", "TC": "TC"}


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def catchment_from_golden(name):
    """Re-create the Catchment the golden run used and check the generator still produces the stored inputs."""
    z, meta = load_golden(name)
    kw = dict(meta["kwargs"])
    if "report_opts" in kw:
        kw["report_opts"] = {k: tuple(v) for k, v in kw["report_opts"].items()}
    if "start" in kw:
        kw["start"] = datetime.date(*kw["start"])
    cat = synth.make_catchment(nsteps=meta["nsteps"], **kw)
    for k, v in cat.static_maps().items():
        assert np.array_equal(z["map/" + k], v), f"synthetic generator drifted for {name}:{k}"
    return cat, z, meta


def golden_forcing(z, t):
    """Stored forcing rasters of model step t (1-based)."""
    out = {k[len("forcing/"):]: z[k][t - 1] for k in z.files if k.startswith("forcing/")}
    for k in ("kc", "seep", "llev"):                                   # map series with gaps: an all-NaN map stands for "no file"
        if k in out and np.isnan(out[k]).all():
            del out[k]
    return out


def bit_equal(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return (a.view(np.int32) == b.view(np.int32)) | (a == b) | (np.isnan(a) & np.isnan(b))
