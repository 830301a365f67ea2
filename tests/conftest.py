import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
for p in (ROOT, REF):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_pyhf: needs the staged reference (baseline/_ref)")


def have_pyhf():
    if not os.path.isdir(os.path.join(REF, "pyhf")) and os.path.isdir("/root/reference/src/pyhf"):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import make_ref

        make_ref.main()
    try:
        import pyhf  # noqa: F401

        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    ok = have_pyhf()
    skip = pytest.mark.skip(reason="reference pyhf not staged (baseline/_ref)")
    for item in items:
        if "needs_pyhf" in item.keywords and not ok:
            item.add_marker(skip)


class Golden:
    """one tests/golden/*.npz file: reference inputs/outputs + the plan tables"""

    def __init__(self, path):
        self.path = path
        self.name = os.path.basename(path)[:-4]
        self.z = np.load(path, allow_pickle=False)
        self.meta = json.loads(str(self.z["meta_json"]))

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z.files

    @property
    def has_plan(self):
        return "plan__meta" in self.z.files

    def plan_arrays(self):
        return {k[6:]: self.z[k] for k in self.z.files if k.startswith("plan__")}

    def host_plan(self):
        from pyhf_b200.plan import HostPlan

        if self.has_plan:
            return HostPlan.from_arrays(self.plan_arrays())
        return synthetic_host_plan(self.meta["generator"])

    def oracle_plan(self):
        from oracle import hf_oracle as O

        if self.has_plan:
            return O.Plan(self.plan_arrays())
        return O.Plan(synthetic_host_plan(self.meta["generator"]).arrays())

    def spec(self):
        if "spec_json" in self.z.files:
            return json.loads(str(self.z["spec_json"]))
        from pyhf_b200 import synth

        return getattr(synth, self.meta["generator"])()

    def model(self):
        import pyhf

        kw = dict(self.meta.get("settings") or {})
        return pyhf.Model(self.spec(), poi_name=self.meta.get("poi_name", "mu"), validate=False, **kw)


_synth_cache = {}


def synthetic_host_plan(generator):
    """plan of a synthetic workspace: through the live reference model when pyhf is staged, else
    through the pyhf-free compiler (the two are asserted equal in tests/test_compile.py)"""
    if generator not in _synth_cache:
        from pyhf_b200 import synth

        spec = getattr(synth, generator)()
        try:
            import pyhf

            from pyhf_b200.plan import extract_plan

            model = pyhf.Model(spec, poi_name="mu", validate=False)
            _synth_cache[generator] = (extract_plan(model), model)
        except ImportError:
            from pyhf_b200.compile import compile_spec

            _synth_cache[generator] = (compile_spec(spec, poi_name="mu"), None)
    return _synth_cache[generator][0]


def golden_names(with_plan_only=False):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        n = os.path.basename(p)[:-4]
        if n in ("hello_toys", "scan_2bin_2channel", "teststats", "c4_fits"):
            continue
        out.append(n)
    return out


SMALL = [n for n in golden_names() if n not in ("c3", "c4", "c4_fits")]
# models built with clip_sample_data / clip_bin_data (pdf.py:696-709): the clipped likelihood has
# kinks and flat stretches, so fits are held to "never a worse minimum than the reference" only
CLIPPED = {"clip_sample0", "clip_bin0", "clip_both", "mixed_clip"}
FIT_GOLDENS = [n for n in SMALL if n != "mixed_clip"]  # goldens that carry reference fits


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = Golden(os.path.join(GOLDEN, f"{name}.npz"))
        return cache[name]

    return load


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyhf_b200 import _lib

    _lib.require()  # a missing native library must fail the GPU tests loudly, not skip them
    return torch.device("cuda", 0)
