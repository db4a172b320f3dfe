"""Layout-exact CPU model of the slab kernels behind ``mlsim_slab_*`` (TEST INFRASTRUCTURE ONLY).

Every method consumes and produces the SAME buffers, in the SAME layouts, as the CUDA entry point of the same
name (include/mlsim.h "slab decomposition"), with the arithmetic taken from the fp64 oracle:

  stage_a  <-> k_explicit_rk<.., SLAB> + k_div_rfft<.., SLAB>   (rows [-1, nxl) from halo-extended buffers, forward
                                                                 spectral chunks [world][batch][nxl][kywp])
  stage_b  <-> k_xsplit_fwd + k_col_solve + k_xsplit_inv        (four-step x-transform with outer radix r0, backward
                                                                 chunks [world][batch][nxl+1][kywp] with the halo row)
  stage_c  <-> k_irfft_grad<.., SLAB>
  cnn_correct <-> conv kernels on rows [lo, hi) with zero padding at the artificial edges

so ``SlabStepper`` + ``SlabComm`` (the product's sequencing and exchanges) can be run under gloo on CPU and compared
with the undecomposed oracle: halo widths, chunk layouts, the duplicated pressure halo row and the split transform
are all exercised without a GPU.
"""
import math

import numpy as np
import torch

from ml_accelerated_simulation_b200.slab import CNN_HALO, HALO, SlabGeometry
from oracle import cnn as OC
from oracle import solver as S


class ModelSlabBackend:
    def __init__(self, geom: SlabGeometry, cfg: S.SolverConfig, r0=None, cnn=None):
        assert cfg.nx == geom.nx and cfg.ny == geom.ny
        self.geom = geom
        self.cfg = cfg
        self.dt = cfg.dt
        self.r0 = geom.r0 if r0 is None else r0
        self.cnn = cnn                               # (configs, state_dict) or None
        g = geom
        inv = S.inverse_eigenvalues(cfg, torch.float64)                 # [nx, nyc]
        self.inv_local = (inv[:, g.ky0:g.ky0 + g.kyv] / (g.nx * g.ny)).numpy()   # 1/(nx*ny) folded in, as in plan.cu

    # ------------------------------------------------------------------ buffers
    def new_field_pair(self):
        return torch.zeros(self.geom.field_shape(), dtype=torch.float64)

    def new_spec(self, backward):
        return torch.zeros(self.geom.spec_shape(backward), dtype=torch.float64)

    def new_pressure(self):
        g = self.geom
        return torch.zeros((g.batch, g.nxl, g.ny), dtype=torch.float64)

    # ------------------------------------------------------------------ helpers
    def _div_rfft(self, out, spec_send):
        g, c = self.geom, self.cfg
        H, n = HALO, g.nxl
        u = out[0, :, H - 1:H + n]                    # rows -1 .. nxl-1
        v = out[1, :, H:H + n]
        div = (u[:, 1:] - u[:, :-1]) / c.hx + (v - torch.roll(v, 1, -1)) / c.hy
        spec = torch.fft.rfft(div, dim=-1) * 1.0      # unnormalised forward, [B, nxl, nyc]
        spec_send.zero_()
        for r in range(g.world):
            k0 = r * g.kyw
            k1 = min(k0 + g.kyw, g.nyc)
            blk = torch.view_as_real(spec[:, :, k0:k1].contiguous())     # [B, nxl, kv, 2]
            spec_send[r, :, :, :2 * (k1 - k0)] = blk.reshape(g.batch, n, -1)

    # ------------------------------------------------------------------ entry points
    def stage_a(self, stage, us, u0, acc, out, spec_send):
        g, c = self.geom, self.cfg
        H, n, dt = HALO, g.nxl, self.dt
        lo, hi = H - 3, H + n + 2                      # rows -3 .. nxl+1 feed outputs -1 .. nxl-1
        ku, kv = S.explicit_terms(us[0, :, lo:hi], us[1, :, lo:hi], c, dt)
        ku, kv = ku[:, 2:-2], kv[:, 2:-2]             # rows -1 .. nxl-1 (the roll wrap only touches the cut rows)
        rows = slice(H - 1, H + n)
        ca, cs = ((dt / 6, 0.5 * dt), (dt / 3, 0.5 * dt), (dt / 3, dt), (0.0, dt / 6))[stage]
        for comp, k in enumerate((ku, kv)):
            if stage == 0:
                acc[comp, :, rows] = u0[comp, :, rows] + ca * k
                new = u0[comp, :, rows] + cs * k
            elif stage < 3:
                acc[comp, :, rows] = acc[comp, :, rows] + ca * k
                new = u0[comp, :, rows] + cs * k
            else:
                new = acc[comp, :, rows] + cs * k
            out[comp, :, rows] = new
        self._div_rfft(out, spec_send)

    def divergence_rfft(self, uv, spec_send):
        self._div_rfft(uv, spec_send)

    def stage_b(self, spec_recv, spec_send2):
        g = self.geom
        n, nx, kv, r0 = g.nxl, g.nx, g.kyv, self.r0
        n2 = nx // r0
        # gather [B, nx, kyv]: x = r*nxl + xl
        rec = spec_recv[..., :2 * kv].reshape(g.world, g.batch, n, kv, 2)
        full = torch.view_as_complex(rec.permute(1, 0, 2, 3, 4).reshape(g.batch, nx, kv, 2).contiguous()).numpy()
        # pass A: x = n2len*n1 + n2 ; T[b][k1][n2] = w^(n2*k1) * DFT_r0 over n1
        a = full.reshape(g.batch, r0, n2, kv)
        t = np.fft.fft(a, axis=1)
        w = np.exp(-2j * np.pi * np.outer(np.arange(r0), np.arange(n2)) / nx)        # [k1, n2]
        t = t * w[None, :, :, None]
        # K3 on every (b, k1) image: FFT over n2 -> k2, kx = k1 + r0*k2, scale, inverse FFT (unnormalised)
        th = np.fft.fft(t, axis=2)
        kx = np.arange(r0)[:, None] + r0 * np.arange(n2)[None, :]                    # [k1, k2]
        th = th * self.inv_local[kx][None]                                           # [B, k1, k2, kv]
        t = np.fft.ifft(th, axis=2) * n2
        # pass C: conj twiddle, inverse DFT_r0 over k1 -> n1
        t = t * np.conj(w)[None, :, :, None]
        p = (np.fft.ifft(t, axis=1) * r0).reshape(g.batch, nx, kv)
        ph = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(p)))           # [B, nx, kv, 2]
        spec_send2.zero_()
        for r in range(g.world):
            rows = [(r * n + xl) % nx for xl in range(n + 1)]                        # row nxl = halo (next slab's row 0)
            spec_send2[r, :, :, :2 * kv] = ph[:, rows].reshape(g.batch, n + 1, -1)

    def stage_c(self, spec_recv2, uv, p=None):
        g, c = self.geom, self.cfg
        H, n = HALO, g.nxl
        ph = torch.zeros((g.batch, n + 1, g.nyc), dtype=torch.complex128)
        for r in range(g.world):
            k0 = r * g.kyw
            k1 = min(k0 + g.kyw, g.nyc)
            blk = spec_recv2[r, :, :, :2 * (k1 - k0)].reshape(g.batch, n + 1, k1 - k0, 2).contiguous()
            ph[:, :, k0:k1] = torch.view_as_complex(blk)
        pr = torch.fft.irfft(ph, n=g.ny, dim=-1) * g.ny                              # unnormalised c2r
        p0 = pr[:, :n]
        uv[0, :, H:H + n] -= (pr[:, 1:] - p0) / c.hx
        uv[1, :, H:H + n] -= (torch.roll(p0, -1, -1) - p0) / c.hy
        if p is not None:
            p.copy_(p0)

    def cnn_correct(self, uv, scale):
        g = self.geom
        H, n = HALO, g.nxl
        lo = H if g.rank == 0 else H - CNN_HALO
        hi = H + n if g.rank == g.world - 1 else H + n + CNN_HALO
        configs, sd = self.cnn
        u, v = OC.correction_add(uv[0, :, lo:hi], uv[1, :, lo:hi], configs, sd, scale)
        uv[0, :, lo:hi] = u
        uv[1, :, lo:hi] = v
