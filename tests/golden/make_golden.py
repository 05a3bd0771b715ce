"""Generates tests/golden/*.npz|json by running the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py

For every fixture the reference modules (imported from /root/reference through
oracle/ref_shim.py) and the CPU oracle (oracle/denoiser_ref.py, oracle/cond_ref.py) are run on
identical weights and inputs; the script asserts they agree to fp32 round-off and stores the
REFERENCE's outputs.  That pins the oracle; the GPU tests then compare the CUDA path with the
oracle and with these fixtures.  Weights and inputs are not stored: they are regenerated from
seeds by `hifidiff_b200.testing.random_state` (keyed by tensor name) and `inputs()` below.

The reference contains no golden vectors of its own, so these are the pins.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cond_ref, denoiser_ref, ref_shim, schedulers_ref  # noqa: E402
from hifidiff_b200 import testing  # noqa: E402
import hifidiff_b200 as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LEVELS = [(128, 16), (256, 8), (512, 4), (1024, 2), (2048, 1)]


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def inputs(kind: str, batch: int, seed: int = 0):
    """Deterministic inputs shared by make_golden.py and the tests."""
    if kind == "latents":
        return torch.randn((batch, 4, 16, 16), generator=gen(100 + seed))
    if kind == "cr_face":
        return torch.rand((batch, 3, 128, 128), generator=gen(200 + seed))
    if kind == "cr_latent":
        return torch.randn((batch, 4, 16, 16), generator=gen(300 + seed))
    raise KeyError(kind)


TRAJ_EPS_GAIN = 0.15  # see hifidiff_b200.testing.random_state


def load_random(module, seed, eps_gain=1.0):
    sd = module.state_dict()
    new = testing.random_state({k: v.shape for k, v in sd.items()}, {k: v.dtype for k, v in sd.items()}, seed, eps_gain)
    module.load_state_dict(new)
    module.eval()
    return new


def tap_stats(taps):
    out = {}
    for k, v in taps.items():
        v = v.float()
        out[k] = [float(v.mean()), float(v.std()), float(v.norm())]
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load()
    summary = {}
    t0 = time.time()

    # ---- 1. state_dict layout + same-seed default init -------------------------------------------
    layout = {}
    for name, ref_cls, my_cls, args in (("Denoiser", ref.Denoiser, H.Denoiser, (16,)),
                                        ("FusedDenoiser", ref.FusedDenoiser, H.FusedDenoiser, (16,)),
                                        ("FacialRefiner", ref.FacialRefiner, H.FacialRefiner, ())):
        torch.manual_seed(0)
        r = ref_cls(*args)
        torch.manual_seed(0)
        m = my_cls(*args)
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys()), f"{name}: key order differs"
        for k in rs:
            assert rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype, (name, k)
            assert torch.equal(rs[k], ms[k]), f"{name}: default init of {k} differs under the same seed"
        layout[name] = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in rs.items()]
        print(f"[layout] {name}: {len(rs)} entries identical (keys, shapes, dtypes, seeded init)")
        del r, m
    with open(os.path.join(OUT, "state_dict_layout.json"), "w") as f:
        json.dump(layout, f)

    # ---- 2. one ConditionalNAFBlock per level ----------------------------------------------------
    blocks = {}
    for lvl, (c, n) in enumerate(LEVELS):
        blk = ref.ConditionalNAFBlock(c, 512)
        sd = load_random(blk, seed=10 + lvl)
        x = torch.randn((2, c, n, n), generator=gen(400 + lvl))
        temb = torch.randn((2, 512), generator=gen(500 + lvl))
        with torch.no_grad():
            y_ref, _ = blk([x, temb])
            y_orc = denoiser_ref.cond_naf_block(sd, "", x, temb)
        e = rel_l2(y_orc, y_ref)
        assert e < 2e-6, (lvl, e)
        blocks[f"level{lvl}"] = y_ref.numpy()
        summary[f"block_level{lvl}_oracle_vs_ref"] = e
        print(f"[block] level {lvl} c={c} n={n}: oracle vs reference rel-L2 {e:.2e}")
    np.savez_compressed(os.path.join(OUT, "naf_blocks.npz"), **blocks)

    # ---- 3. one HybridCrossAttention per level ---------------------------------------------------
    hcas = {}
    for j, (c, n) in enumerate(LEVELS[::-1]):
        mod = ref.HybridCrossAttention(c)
        sd = load_random(mod, seed=20 + j)
        f_g = torch.randn((2, c, n, n), generator=gen(600 + j))
        f_d = torch.randn((2, c, n, n), generator=gen(700 + j))
        with torch.no_grad():
            y_ref = mod(f_g, f_d)
            y_orc = denoiser_ref.hca(sd, "", f_g, f_d)
        e = rel_l2(y_orc, y_ref)
        assert e < 2e-6, (j, e)
        hcas[f"hca{j}"] = y_ref.numpy()
        summary[f"hca{j}_oracle_vs_ref"] = e
        print(f"[hca] {j} dim={c} n={n}: oracle vs reference rel-L2 {e:.2e}")
    np.savez_compressed(os.path.join(OUT, "hca.npz"), **hcas)

    # ---- 4. Denoiser: one step, B=2, per-face timesteps; timestep argument forms ------------------
    den = ref.Denoiser(16)
    sd = load_random(den, seed=1)
    x = inputs("latents", 2)
    t = torch.tensor([500, 37], dtype=torch.long)
    with torch.no_grad():
        y_ref = den(x, t).sample
        taps = {}
        y_orc = denoiser_ref.denoiser_forward(sd, x, t, taps)
        e = rel_l2(y_orc, y_ref)
        assert e < 5e-6, e
        forms = {}
        for label, tv in (("int", 500), ("float", 500.0), ("zero_d", torch.tensor(500)), ("len1", torch.tensor([500.0])),
                          ("long_B", torch.tensor([500, 500]))):
            if label == "len1":
                continue  # Denoiser does not broadcast a length-1 tensor (model.py:107-108); FusedDenoiser does
            yr = den(x, tv).sample
            yo = denoiser_ref.denoiser_forward(sd, x, tv)
            assert rel_l2(yo, yr) < 5e-6, label
            forms[label] = yr.numpy()
    summary["denoiser_oracle_vs_ref"] = e
    print(f"[denoiser] oracle vs reference rel-L2 {e:.2e}; |eps| rms {float(y_ref.pow(2).mean().sqrt()):.3f}")
    np.savez_compressed(os.path.join(OUT, "denoiser_step.npz"), eps=y_ref.numpy(), t=t.numpy(),
                        eps_t500=forms["int"], tap_names=np.array(list(taps.keys())),
                        tap_stats=np.array([tap_stats(taps)[k] for k in taps], dtype=np.float64),
                        tap_intro=taps["intro"].numpy(), tap_mid7=taps["middle_blks.7"].numpy(),
                        tap_dec31=taps["decoders.3.1"].numpy(), tap_time=taps["time_mlp"].numpy())

    # DDIM-50 trajectory, B=1, reference model driven by the oracle scheduler (diffusers is absent)
    load_random(den, seed=1, eps_gain=TRAJ_EPS_GAIN)
    sched = schedulers_ref.DDIMSchedulerRef(clip_sample=False)
    xT = inputs("latents", 1, seed=7)
    with torch.no_grad():
        traj = []
        x0 = schedulers_ref.sample_loop(lambda xx, tt: den(xx, torch.full((1,), tt, dtype=torch.long)).sample, xT, sched, 50,
                                        on_step=lambda i, tt, eps, xn: traj.append(xn.clone()) if i in (0, 9, 24, 49) else None)
    np.savez_compressed(os.path.join(OUT, "denoiser_ddim50.npz"), x0=x0.numpy(), x_after_step0=traj[0].numpy(),
                        x_after_step9=traj[1].numpy(), x_after_step24=traj[2].numpy())
    print(f"[denoiser] DDIM-50 x0 rms {float(x0.pow(2).mean().sqrt()):.3f} range [{float(x0.min()):.2f},{float(x0.max()):.2f}]")
    del den, sd

    # ---- 5. FusedDenoiser: one step with synthetic priors / identity ------------------------------
    fus = ref.FusedDenoiser(16)
    sd = load_random(fus, seed=2)
    x = inputs("latents", 2, seed=1)
    priors, ident = testing.synthetic_condition(2, 16, seed=0)
    t = torch.tensor([980, 3], dtype=torch.long)
    with torch.no_grad():
        y_ref = fus(x, t, priors, ident).sample
        taps = {}
        y_orc = denoiser_ref.fused_denoiser_forward(sd, x, t, priors, ident, taps)
        e = rel_l2(y_orc, y_ref)
        assert e < 5e-6, e
        for label, tv in (("int", 980), ("float", 980.0), ("zero_d", torch.tensor(980)), ("len1", torch.tensor([980])),
                          ("float_B", torch.tensor([980.0, 3.0]))):
            yr = fus(x, tv, priors, ident).sample
            yo = denoiser_ref.fused_denoiser_forward(sd, x, tv, priors, ident)
            assert rel_l2(yo, yr) < 5e-6, label
    summary["fused_oracle_vs_ref"] = e
    print(f"[fused] oracle vs reference rel-L2 {e:.2e}; |eps| rms {float(y_ref.pow(2).mean().sqrt()):.3f}")
    np.savez_compressed(os.path.join(OUT, "fused_step.npz"), eps=y_ref.numpy(), t=t.numpy(),
                        tap_names=np.array(list(taps.keys())),
                        tap_stats=np.array([tap_stats(taps)[k] for k in taps], dtype=np.float64),
                        tap_hca0=taps["hcas.0"].numpy(), tap_hca4=taps["hcas.4"].numpy(),
                        tap_up0=taps["ups.0"].numpy())
    load_random(fus, seed=2, eps_gain=TRAJ_EPS_GAIN)
    sched = schedulers_ref.DDIMSchedulerRef(clip_sample=False)
    xT = inputs("latents", 1, seed=8)
    p1, i1 = testing.synthetic_condition(1, 16, seed=3)
    with torch.no_grad():
        x0 = schedulers_ref.sample_loop(lambda xx, tt: fus(xx, torch.full((1,), tt, dtype=torch.long), p1, i1).sample,
                                        xT, sched, 50)
    np.savez_compressed(os.path.join(OUT, "fused_ddim50.npz"), x0=x0.numpy())
    print(f"[fused] DDIM-50 x0 rms {float(x0.pow(2).mean().sqrt()):.3f}")
    del fus, sd

    # ---- 6. FacialRefiner: FPG + IDC + FusedDenoiser ---------------------------------------------
    refm = ref.FacialRefiner()
    sd = load_random(refm, seed=3)
    x = inputs("latents", 1, seed=2)
    cr_face, cr_latent = inputs("cr_face", 1), inputs("cr_latent", 1)
    with torch.no_grad():
        y_ref = refm(x, torch.tensor([640]), cr_face, cr_latent).sample
        pri_ref = refm.fpg(cr_latent)
        id_ref = refm.idc(cr_face)
        taps = {}
        y_orc = cond_ref.refiner_forward(sd, x, torch.tensor([640]), cr_face, cr_latent, taps)
    e = rel_l2(y_orc, y_ref)
    assert e < 5e-6, e
    for j in range(5):
        assert rel_l2(taps[f"prior{j}"], pri_ref[j]) < 5e-6, j
    assert rel_l2(taps["identity"], id_ref) < 5e-6
    summary["refiner_oracle_vs_ref"] = e
    print(f"[refiner] oracle vs reference rel-L2 {e:.2e}")
    np.savez_compressed(os.path.join(OUT, "refiner_step.npz"), eps=y_ref.numpy(), identity=id_ref.numpy(),
                        **{f"prior{j}": pri_ref[j].numpy() for j in range(5)})

    summary["seconds"] = time.time() - t0
    summary["torch"] = torch.__version__
    with open(os.path.join(OUT, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()
