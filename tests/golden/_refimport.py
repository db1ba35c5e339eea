"""Import the UNMODIFIED reference PIPS modules from /root/reference (build container only).

`import sam_pt.point_tracker.pips` fails as-is because sam_pt/point_tracker/__init__.py:2-7 eagerly pulls
superglue/tapir/cotracker (matplotlib, tensorflow, ... absent).  Work-around: register empty
namespace packages whose __path__ points at the reference dirs, then import the leaf modules.
"""
import importlib
import sys
import types

REF = "/root/reference"


def import_reference_pips():
    saved = {k: v for k, v in sys.modules.items() if k == "sam_pt" or k.startswith("sam_pt.")}
    for k in saved:
        del sys.modules[k]
    pkgs = {
        "sam_pt": f"{REF}/sam_pt",
        "sam_pt.point_tracker": f"{REF}/sam_pt/point_tracker",
    }
    for name, path in pkgs.items():
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    sys.modules["sam_pt"].point_tracker = sys.modules["sam_pt.point_tracker"]
    for leaf in ("utils.basic", "utils.samp", "utils.misc", "utils.saverloader"):
        importlib.import_module("sam_pt.point_tracker." + leaf)
    tracker = importlib.import_module("sam_pt.point_tracker.tracker")
    pips = importlib.import_module("sam_pt.point_tracker.pips.pips")
    ptracker = importlib.import_module("sam_pt.point_tracker.pips.tracker")
    out = {"Pips": pips.Pips, "PipsPointTracker": ptracker.PipsPointTracker, "PointTracker": tracker.PointTracker,
           "pips_module": pips}
    # restore whatever `sam_pt` was importable before (the product package, if on sys.path)
    for k in [k for k in sys.modules if k == "sam_pt" or k.startswith("sam_pt.")]:
        del sys.modules[k]
    sys.modules.update(saved)
    return out
