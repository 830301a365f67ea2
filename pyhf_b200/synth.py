"""Synthetic HistFactory workspaces of the shapes named in BASELINE.json.

These are plain JSON-like specs in pyhf's model-spec format (channels / samples / modifiers);
they are *inputs* to the path, not part of it.  ``numpy.random.default_rng(seed)`` draws happen in
exactly the nested order documented in SURVEY.md Appendix C so that every run (here, on the GPU
box, in the golden-vector script) sees the same workspace.
"""
import numpy as np

__all__ = ["synth_spec", "spec_c3", "spec_c4", "spec_hello", "param_batch"]


def synth_spec(
    n_channels,
    bins_per_channel,
    n_bkg,
    n_histosys,
    n_normsys,
    staterror_channels,
    n_shapesys_channels=0,
    seed=1234,
):
    rng = np.random.default_rng(seed)
    nb = bins_per_channel
    channels = []
    for c in range(n_channels):
        sig = rng.uniform(1.0, 10.0, nb).round(6)
        samples = [
            {
                "name": "signal",
                "data": sig.tolist(),
                "modifiers": [{"name": "mu", "type": "normfactor", "data": None}],
            }
        ]
        for s in range(n_bkg):
            nom = rng.uniform(20.0, 200.0, nb).round(6)
            mods = []
            for k in range(n_normsys):
                lo = float(np.round(rng.uniform(0.85, 0.98), 6))
                hi = float(np.round(rng.uniform(1.02, 1.15), 6))
                mods.append(
                    {"name": f"ns_{k:03d}", "type": "normsys", "data": {"hi": hi, "lo": lo}}
                )
            for k in range(n_histosys):
                up = (nom * (1.0 + rng.uniform(0.0, 0.08, nb))).round(6)
                dn = (nom * (1.0 - rng.uniform(0.0, 0.08, nb))).round(6)
                mods.append(
                    {
                        "name": f"hs_{k:03d}",
                        "type": "histosys",
                        "data": {"hi_data": up.tolist(), "lo_data": dn.tolist()},
                    }
                )
            if c < staterror_channels:
                mods.append(
                    {
                        "name": f"staterror_ch{c:02d}",
                        "type": "staterror",
                        "data": (0.5 * np.sqrt(nom)).round(6).tolist(),
                    }
                )
            if s == 0 and c >= n_channels - n_shapesys_channels:
                mods.append(
                    {
                        "name": f"shapesys_ch{c:02d}",
                        "type": "shapesys",
                        "data": (0.1 * nom).round(6).tolist(),
                    }
                )
            samples.append({"name": f"bkg{s}", "data": nom.tolist(), "modifiers": mods})
        channels.append({"name": f"ch{c:02d}", "samples": samples})
    return {"channels": channels}


def spec_c3(seed=1234):
    """BASELINE.json config 3: 1 channel x 100 bins, 5 samples, 19 normsys + 1 shapesys (P=120)."""
    return synth_spec(1, 100, 4, 0, 19, 0, n_shapesys_channels=1, seed=seed)


def spec_c4(seed=1234):
    """BASELINE.json configs 4/5: 20 channels x 50 bins, 100 histosys + 99 normsys + staterror (P=300)."""
    return synth_spec(20, 50, 4, 100, 99, 2, seed=seed)


def spec_hello():
    """BASELINE.json config 1 == pyhf.simplemodels.uncorrelated_background([12,11],[50,52],[3,7])
    (reference: src/pyhf/simplemodels.py:13-91), written out so no pyhf import is needed."""
    return {
        "channels": [
            {
                "name": "singlechannel",
                "samples": [
                    {
                        "name": "signal",
                        "data": [12.0, 11.0],
                        "modifiers": [{"name": "mu", "type": "normfactor", "data": None}],
                    },
                    {
                        "name": "background",
                        "data": [50.0, 52.0],
                        "modifiers": [
                            {"name": "uncorr_bkguncrt", "type": "shapesys", "data": [3.0, 7.0]}
                        ],
                    },
                ],
            }
        ]
    }


def param_batch(init, lo, hi, is_alpha, is_gamma, poi_index, B, seed=0):
    """Parameter batch of SURVEY.md section 8(d) config 4: alpha~N(0,0.7), gamma~1+N(0,0.05),
    mu~U(0,2), clipped to bounds.  ``is_alpha``/``is_gamma`` are boolean masks over parameters."""
    rng = np.random.default_rng(seed)
    P = len(init)
    x = np.tile(np.asarray(init, dtype=np.float64), (B, 1))
    a = rng.normal(0.0, 0.7, size=(B, P))
    g = 1.0 + rng.normal(0.0, 0.05, size=(B, P))
    m = rng.uniform(0.0, 2.0, size=B)
    x = np.where(np.asarray(is_alpha)[None, :], a, x)
    x = np.where(np.asarray(is_gamma)[None, :], g, x)
    x[:, poi_index] = m
    return np.clip(x, np.asarray(lo)[None, :], np.asarray(hi)[None, :])
