"""Fit the reference DeepSDF `Decoder` (8x256, L=64, latent_in=[4], weight-norm) to an analytic
latent-conditioned ellipsoid SDF, and store its state_dict as a golden fixture.

Why: the trained DeepSDF weights are not in the reference repo (README.md:112) and a random-init
decoder has an almost constant output (no zero level set -> empty render band -> is_good=False),
see SURVEY.md section 8c / Appendix B.1.  The architecture spec is the north-star one
(deep_sdf/deep_sdf_decoder.py:10-72 constructed as in SURVEY.md section 8d).

Run here (CPU, ~2-3 min per decoder):  python tools/fit_fixture_decoder.py cars|chairs
Output: tests/golden/decoder_<name>.npz  (weight_g / weight_v / bias per layer + spec)
"""
import os
import sys
import json
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness  # noqa: E402

SPEC = dict(latent_size=64, dims=[256] * 8, dropout=list(range(8)), dropout_prob=0.2,
            norm_layers=list(range(8)), latent_in=[4], xyz_in_all=False, use_tanh=False,
            latent_dropout=False, weight_norm=True)

CLASSES = {
    "cars": dict(seed=0, radii=(0.30, 0.25, 0.60)),
    "chairs": dict(seed=1, radii=(0.35, 0.45, 0.35)),
}


def ellipsoid_sdf(x, radii):
    """First-order ellipsoid distance  k0 (k0 - 1) / k1  with k0=|x/r|, k1=|x/r^2|."""
    k0 = (x / radii).norm(dim=-1)
    k1 = (x / (radii * radii)).norm(dim=-1).clamp_min(1e-9)
    return k0 * (k0 - 1.0) / k1


def sample_batch(gen, base_radii, n_lat=64, n_pts=64):
    z = 0.15 * torch.randn(n_lat, 64, generator=gen)
    radii = base_radii[None, :] * (1.0 + z[:, :3])                      # (n_lat, 3)
    radii = radii.clamp_min(0.05)
    n_near = int(0.85 * n_pts)
    dirs = torch.randn(n_lat, n_near, 3, generator=gen)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    surf = dirs * radii[:, None, :]
    sigma = torch.where(torch.rand(n_lat, n_near, 1, generator=gen) < 0.5,
                        torch.tensor(0.05), torch.tensor(0.01))
    near = surf + sigma * torch.randn(n_lat, n_near, 3, generator=gen)
    uni = torch.rand(n_lat, n_pts - n_near, 3, generator=gen) * 2.0 - 1.0
    x = torch.cat([near, uni], dim=1)                                   # (n_lat, n_pts, 3)
    tgt = ellipsoid_sdf(x, radii[:, None, :]).clamp(-0.1, 0.1)
    inp = torch.cat([z[:, None, :].expand(-1, n_pts, -1), x], dim=-1)
    return inp.reshape(-1, 67), tgt.reshape(-1, 1)


def fit(name, steps=1200):
    cfg = CLASSES[name]
    ns = ref_harness.load()
    torch.manual_seed(cfg["seed"])
    torch.set_num_threads(os.cpu_count())
    dec = ns.decoder.Decoder(**SPEC)
    dec.train()
    opt = torch.optim.Adam(dec.parameters(), lr=5e-4)
    gen = torch.Generator().manual_seed(1000 + cfg["seed"])
    base = torch.tensor(cfg["radii"])
    for it in range(steps):
        inp, tgt = sample_batch(gen, base)
        pred = dec(inp)
        loss = (pred.clamp(-0.1, 0.1) - tgt).abs().mean() + 0.1 * (pred - tgt).abs().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it % 200 == 0 or it == steps - 1:
            print(f"[{name}] step {it} loss {loss.item():.5f}", flush=True)
    dec.eval()
    # quality probe: SDF on / outside / inside the z=0 surface
    with torch.no_grad():
        d = torch.randn(2000, 3)
        d = d / d.norm(dim=-1, keepdim=True)
        for f in (0.7, 1.0, 1.3):
            x = d * base * f
            y = dec(torch.cat([torch.zeros(2000, 64), x], -1))
            print(f"[{name}] sdf at {f:.1f}x surface: {y.mean().item():+.4f} +- {y.std().item():.4f}")
    out = {k: v.detach().cpu().numpy() for k, v in dec.state_dict().items()}
    out["spec_json"] = np.frombuffer(json.dumps(SPEC).encode(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"decoder_{name}.npz")
    np.savez(path, **out)
    print("wrote", os.path.abspath(path))


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["cars", "chairs"]):
        fit(n)
