"""CPU restatement of the reference's extended-Tofts model and fitting loop.  TEST INFRASTRUCTURE ONLY (like everything under
oracle/): imported by tests/ and never by the product.

Follows /root/reference/pk_fitting.py: population_aif :28-46, extended_tofts_model_batch :193-231, the optimisation of
fit_volume_gpu :288-368 (initial guess, torch.optim.Adam lr 0.005 over the FULL parameter vectors stepped once per
1024-pixel batch, F.mse_loss, clamps after every step).  Pinned by tests/golden/tofts_fit_80x80.npz, generated from the live
reference class by tests/golden/make_golden_tofts.py.
"""
import torch
import torch.nn.functional as F


def population_aif(t, dose=0.1):
    a1, a2, m1, m2 = 3.99, 4.78, 0.144, 0.0111
    return dose * (a1 * torch.exp(-m1 * t) + a2 * torch.exp(-m2 * t))


def extended_tofts_model_batch(t, ktrans, ve, vp, aif=population_aif, dt=0.01):
    """[N] parameters -> [N, T]; time points without an earlier grid point keep 0 (:210, :216-217)."""
    aif_values = aif(t)
    t_conv = torch.arange(0, t[-1].item(), dt, dtype=torch.float32, device=t.device)
    aif_conv = aif(t_conv)
    cols = []
    for i, ti in enumerate(t):
        mask = t_conv < ti
        if not mask.any():
            cols.append(torch.zeros_like(ktrans))
            continue
        tv, av = t_conv[mask], aif_conv[mask]
        e = torch.exp(-ktrans.view(-1, 1) * (ti - tv.view(1, -1)) / ve.view(-1, 1))
        cols.append(vp * aif_values[i] + ktrans * (torch.sum(av.view(1, -1) * e, dim=1) * dt))
    return torch.stack(cols, dim=1)


def fit_pixels(t, pixels, epochs=100, batch_size=1024, lr=0.005):
    """-> (ktrans, ve, vp, per-epoch mean batch loss).  One Adam over the full vectors, one step per batch (:323-345)."""
    n = pixels.shape[0]
    ktrans = torch.full((n,), 0.05, dtype=torch.float32, requires_grad=True)
    ve = torch.full((n,), 0.1, dtype=torch.float32, requires_grad=True)
    vp = torch.full((n,), 0.01, dtype=torch.float32, requires_grad=True)
    opt = torch.optim.Adam([ktrans, ve, vp], lr=lr)
    losses = []
    nb = (n + batch_size - 1) // batch_size
    for _ in range(epochs):
        tot = 0.0
        for b in range(nb):
            s, e = b * batch_size, min((b + 1) * batch_size, n)
            pred = extended_tofts_model_batch(t, ktrans[s:e], ve[s:e], vp[s:e])
            loss = F.mse_loss(pred, pixels[s:e])
            opt.zero_grad()
            loss.backward()
            opt.step()
            with torch.no_grad():
                ktrans.clamp_(0.0, 1.0)
                ve.clamp_(0.001, 0.5)
                vp.clamp_(0.0, 0.2)
            tot += loss.item()
        losses.append(tot / nb)
    return ktrans.detach(), ve.detach(), vp.detach(), torch.tensor(losses)
