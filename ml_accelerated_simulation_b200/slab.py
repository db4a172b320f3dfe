"""One large periodic grid decomposed into x-slabs over several B200s (BASELINE.json configs[4]: 8192 x 8192).

The step is the reference's loop body (src/inference.py:100-102) -- RK4 with a projection after every stage,
then ``v += model(v)`` -- on a grid too large for the per-step budget of one device.  One process per GPU; rank r
owns the global rows ``[r*nxl, (r+1)*nxl)`` (array axis 0 = x; rows stay contiguous in y).  What crosses NVLink:

  * per RK stage, a halo exchange of the stage state: 3 rows from the lower neighbour and 2 from the upper one
    (radius-2 TVD stencil, plus one more row below so that the explicit kernel can also produce row -1 of u*, the
    x-halo the divergence of local row 0 needs -- this saves a second exchange per stage); periodic in x;
  * per projection, two all-to-alls of the half spectrum: the y-transform is local to a slab, the x-transform needs
    all rows of a spectral column, so column chunk r travels to rank r and back (the backward chunks carry one extra
    row, the x-halo of the pressure);
  * before the correction network, 7 rows of state per interior side (one per 3x3 layer; the network pads with zeros
    at the edges of the global grid, so this exchange is NOT periodic) -- the activations themselves never travel.

All arithmetic is in the sm_100a kernels behind ``mlsim_slab_*`` (include/mlsim.h); this module only sequences them
and moves bytes with ``torch.distributed`` (NCCL on GPUs).  With ``SlabComm(..., p2p=True)`` the two spectral
all-to-alls per projection disappear: the receive buffers are symmetric memory (``torch.distributed._symmetric_memory``,
NVLink peer mappings), the producing kernels (divergence + row FFT; the inverse split pass of the x-solve) store every
column chunk straight into the consuming rank's buffer while they compute, and only a device-side barrier remains.  ``SlabStepper`` is written against a small backend
interface so that the sequencing and every buffer layout can be checked on CPU (gloo, world size 2) against a layout
-exact numpy model in ``tests/slab_model.py``; the product backend is ``CudaSlabBackend`` and there is no other.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .plan import Plan, PlanConfig

HALO = 8          # halo rows on each side of a field buffer (MLSIM_SLAB_HALO)
CNN_HALO = 7      # rows of state the 7-layer 3x3 network needs from each interior neighbour
STAGE_HALO_LO = 3
STAGE_HALO_HI = 2
MAX_COL_FFT = 2048   # longest x-transform done inside shared memory by one column kernel
SPLIT_N2 = 1024      # grids longer than that: nx = r0 * SPLIT_N2 (MLSIM_SPLIT_N2 in csrc/plan.cu)


@dataclasses.dataclass(frozen=True)
class SlabGeometry:
    nx: int            # global rows
    ny: int
    batch: int
    rank: int
    world: int

    def __post_init__(self):
        W = self.world
        if W < 1 or W & (W - 1):
            raise ValueError("world must be a power of two")
        if not 0 <= self.rank < W:
            raise ValueError("rank out of range")
        if self.nx % W or self.nx // W < 32:
            raise ValueError("a slab needs at least 32 rows per rank")
        if self.ny // 2 < W * W:
            raise ValueError("ny too small for this many ranks")

    @property
    def nxl(self) -> int:
        return self.nx // self.world

    @property
    def ext_rows(self) -> int:
        return self.nxl + 2 * HALO

    @property
    def nyc(self) -> int:
        return self.ny // 2 + 1

    @property
    def kyw(self) -> int:          # spectral columns per rank chunk
        return -(-self.nyc // self.world)

    @property
    def kywp(self) -> int:         # padded pitch
        return (self.kyw + 3) & ~3

    @property
    def ky0(self) -> int:
        return self.rank * self.kyw

    @property
    def kyv(self) -> int:          # valid columns of this rank
        return min(self.kyw, self.nyc - self.ky0)

    @property
    def r0(self) -> int:           # outer radix of the split x-transform
        return self.nx // SPLIT_N2 if self.nx > MAX_COL_FFT else 1

    def field_shape(self):
        return (2, self.batch, self.ext_rows, self.ny)

    def spec_shape(self, backward: bool):
        return (self.world, self.batch, self.nxl + (1 if backward else 0), 2 * self.kywp)    # float32 pairs (re, im)


class SlabComm:
    """The three exchanges of the slab step over a torch.distributed process group (None: single rank).

    p2p=True (GPUs of one NVLink domain, world > 1): spectral chunks travel as peer-memory stores issued by the
    producing kernels; ``spectral_exchange`` is then only a cross-rank barrier."""

    BARRIER_TIMEOUT_MS = 20000

    def __init__(self, geom: SlabGeometry, group=None, p2p: bool = False):
        self.g = geom
        self.group = group
        if geom.world > 1 and not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("slab world > 1 needs an initialised torch.distributed process group")
        self.p2p = bool(p2p) and geom.world > 1
        self._stage = {}
        self._symm = []           # (tensor, handle) of the symmetric receive buffers

    # ------------------------------------------------------------------ spectral buffers
    def make_spectral_buffers(self, backend):
        """(send, recv, send2, recv2).  NCCL: four local buffers.  p2p: recv / recv2 are symmetric memory registered
        with the backend (mlsim_slab_set_peers); send / send2 are None (the kernels write into the peers)."""
        if not self.p2p:
            return backend.new_spec(False), backend.new_spec(False), backend.new_spec(True), backend.new_spec(True)
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        bufs, ptrs = [], []
        for backward in (False, True):
            t = symm_mem.empty(self.g.spec_shape(backward), dtype=torch.float32, device=backend.device)
            t.zero_()                                    # padding columns are never written by the kernels
            h = symm_mem.rendezvous(t, grp)
            self._symm.append((t, h))
            bufs.append(t)
            ptrs.append([int(x) for x in h.buffer_ptrs])
        torch.cuda.synchronize(backend.device)
        dist.barrier(group=self.group)                   # every rank's buffers are zeroed before anyone stores into them
        backend.set_peers(ptrs[0], ptrs[1])
        return None, bufs[0], None, bufs[1]

    # ------------------------------------------------------------------ field buffers
    def make_field_buffers(self, backend):
        """(state, w1, w2, acc).  p2p: the three buffers whose halos are exchanged are symmetric memory, so a rank can
        push its edge rows straight into its neighbours' halo rows."""
        if not self.p2p:
            return tuple(backend.new_field_pair() for _ in range(4))
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        out = []
        self._field_peers = {}
        self._backend = backend
        for _ in range(3):
            t = symm_mem.empty(self.g.field_shape(), dtype=torch.float32, device=backend.device)
            t.zero_()
            h = symm_mem.rendezvous(t, grp)
            self._field_peers[t.data_ptr()] = ([int(x) for x in h.buffer_ptrs], h)
            out.append(t)
        torch.cuda.synchronize(backend.device)
        dist.barrier(group=self.group)
        return out[0], out[1], out[2], backend.new_field_pair()

    def fence(self) -> None:
        """Cross-rank barrier on the compute stream (p2p mode: nobody may push into a neighbour that is still working on
        the rows being replaced -- used after the correction network)."""
        if self.p2p:
            self._symm[0][1].barrier(channel=2, timeout_ms=self.BARRIER_TIMEOUT_MS)

    def spectral_exchange(self, send, recv, which: int) -> None:
        """Chunk r of this rank's half spectrum -> rank r (which: 0 forward, 1 backward)."""
        if self.p2p:
            self._symm[which][1].barrier(channel=which, timeout_ms=self.BARRIER_TIMEOUT_MS)
        else:
            self.all_to_all(send, recv)

    def _bufs(self, like: torch.Tensor, n: int):
        key = (n, like.device, like.dtype)
        if key not in self._stage:
            shape = (2, self.g.batch, n, self.g.ny)
            self._stage[key] = tuple(torch.empty(shape, dtype=like.dtype, device=like.device) for _ in range(4))
        return self._stage[key]

    def halo_exchange(self, uv: torch.Tensor, lo: int, hi: int, periodic: bool) -> None:
        """Fill halo rows [-lo, 0) from the lower neighbour's top rows and [nxl, nxl+hi) from the upper neighbour's
        bottom rows.  uv: [2, batch, ext_rows, ny]."""
        g = self.g
        H, n = HALO, g.nxl
        if lo > H or hi > H:
            raise ValueError("halo wider than the buffer margin")
        if g.world == 1:
            if periodic:
                uv[:, :, H - lo:H].copy_(uv[:, :, H + n - lo:H + n])
                uv[:, :, H + n:H + n + hi].copy_(uv[:, :, H:H + hi])
            return
        up, down = (g.rank + 1) % g.world, (g.rank - 1) % g.world
        has_up = periodic or g.rank < g.world - 1
        has_down = periodic or g.rank > 0
        peers = getattr(self, "_field_peers", {}).get(uv.data_ptr()) if self.p2p else None
        if peers is not None:
            # peer memory: one small kernel stores the edge rows into the neighbours' halos, then a device barrier
            ptrs, h = peers
            self._backend.halo_push(uv, ptrs[up] if has_up else None, ptrs[down] if has_down else None, lo, hi)
            h.barrier(channel=0, timeout_ms=self.BARRIER_TIMEOUT_MS)
            return
        s_up, r_dn, _, _ = self._bufs(uv, lo)
        _, _, s_dn, r_up = self._bufs(uv, hi)
        ops = []
        # posting order is the same on every rank (send up, send down, recv from below, recv from above), so the two
        # messages between the same pair of ranks at world == 2 match in order
        if has_up:
            s_up.copy_(uv[:, :, H + n - lo:H + n])
            ops.append(dist.P2POp(dist.isend, s_up, up, self.group))
        if has_down:
            s_dn.copy_(uv[:, :, H:H + hi])
            ops.append(dist.P2POp(dist.isend, s_dn, down, self.group))
        if has_down:
            ops.append(dist.P2POp(dist.irecv, r_dn, down, self.group))
        if has_up:
            ops.append(dist.P2POp(dist.irecv, r_up, up, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if has_down:
            uv[:, :, H - lo:H].copy_(r_dn)
        if has_up:
            uv[:, :, H + n:H + n + hi].copy_(r_up)

    def all_to_all(self, send: torch.Tensor, recv: torch.Tensor) -> None:
        """Chunk r of `send` (dim 0) goes to rank r; chunk r of `recv` comes from rank r."""
        if self.g.world == 1:
            recv.copy_(send)
        else:
            dist.all_to_all_single(recv, send, group=self.group)


class CudaSlabBackend:
    """The product backend: a slab-mode ``mlsim_plan`` (hand-written sm_100a kernels) and torch for memory only."""

    def __init__(self, geom: SlabGeometry, device: int = 0, **physics):
        self.geom = geom
        self.plan = Plan(PlanConfig(nx=geom.nx, ny=geom.ny, batch=geom.batch, device=device, slab_rank=geom.rank,
                                    slab_world=geom.world, **physics))
        self._lib = self.plan._lib
        self._h = self.plan._h
        self.device = self.plan.device
        out = (C.c_int64 * 8)()
        _lib.check(self._lib.mlsim_slab_geometry(self._h, out))
        mine = (geom.nxl, HALO, geom.ext_rows, geom.kyw, geom.kywp, geom.kyv, geom.r0, CNN_HALO)
        if tuple(out) != mine:
            raise RuntimeError(f"slab geometry mismatch between slab.py {mine} and libmlsim {tuple(out)}")

    def new_field_pair(self) -> torch.Tensor:
        return torch.zeros(self.geom.field_shape(), dtype=torch.float32, device=self.device)

    def new_spec(self, backward: bool) -> torch.Tensor:
        return torch.zeros(self.geom.spec_shape(backward), dtype=torch.float32, device=self.device)

    def new_pressure(self) -> torch.Tensor:
        return torch.zeros((self.geom.batch, self.geom.nxl, self.geom.ny), dtype=torch.float32, device=self.device)

    def _st(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()      # None: the kernels store into the registered peer buffers

    def stage_a(self, stage, us, u0, acc, out, spec_send):
        _lib.check(self._lib.mlsim_slab_stage_a(self._h, stage, us.data_ptr(), u0.data_ptr(), acc.data_ptr(),
                                                out.data_ptr(), self._p(spec_send), self._st()))

    def divergence_rfft(self, uv, spec_send):
        _lib.check(self._lib.mlsim_slab_divergence_rfft(self._h, uv.data_ptr(), self._p(spec_send), self._st()))

    def stage_b(self, spec_recv, spec_send2):
        _lib.check(self._lib.mlsim_slab_stage_b(self._h, spec_recv.data_ptr(), self._p(spec_send2), self._st()))

    def stage_c(self, spec_recv2, uv, p=None):
        _lib.check(self._lib.mlsim_slab_stage_c(self._h, spec_recv2.data_ptr(), uv.data_ptr(),
                                                None if p is None else p.data_ptr(), self._st()))

    def cnn_correct(self, uv, scale):
        _lib.check(self._lib.mlsim_slab_cnn_correct(self._h, uv.data_ptr(), float(scale), self._st()))

    def halo_push(self, uv, up_ptr, dn_ptr, lo, hi):
        _lib.check(self._lib.mlsim_slab_halo_push(self._h, uv.data_ptr(), up_ptr, dn_ptr, int(lo), int(hi), self._st()))

    def set_peers(self, fwd_ptrs, bwd_ptrs):
        n = len(fwd_ptrs)
        a = (C.c_uint64 * n)(*fwd_ptrs)
        b = (C.c_uint64 * n)(*bwd_ptrs)
        _lib.check(self._lib.mlsim_slab_set_peers(self._h, a, b))

    def set_cnn(self, layers):
        self.plan.set_cnn(layers)

    def close(self):
        self.plan.close()


class SlabStepper:
    """Sequences one rank's share of the rollout step.  ``state`` is the halo-extended field pair; use
    ``scatter`` / ``gather`` to move between a global [batch, nx, ny] array and the slabs."""

    def __init__(self, backend, comm: Optional[SlabComm] = None):
        self.be = backend
        self.g: SlabGeometry = backend.geom
        self.comm = comm if comm is not None else SlabComm(self.g)
        b = backend
        self.state, self.w1, self.w2, self.acc = self.comm.make_field_buffers(b)
        self.send, self.recv, self.send2, self.recv2 = self.comm.make_spectral_buffers(b)
        self.pressure = None
        self.exchanges = 0

    # ------------------------------------------------------------------ data movement
    def set_local(self, u_local: torch.Tensor, v_local: torch.Tensor) -> None:
        """u_local, v_local: this rank's rows, [batch, nxl, ny]."""
        n = self.g.nxl
        self.state[0, :, HALO:HALO + n].copy_(u_local)
        self.state[1, :, HALO:HALO + n].copy_(v_local)

    def scatter(self, u_global: torch.Tensor, v_global: torch.Tensor) -> None:
        """Every rank passes the same global [batch, nx, ny] arrays (host or device) and keeps its rows."""
        n, r = self.g.nxl, self.g.rank
        self.set_local(u_global[:, r * n:(r + 1) * n], v_global[:, r * n:(r + 1) * n])

    def local(self):
        n = self.g.nxl
        return self.state[0, :, HALO:HALO + n], self.state[1, :, HALO:HALO + n]

    def gather(self):
        """Global (u, v) on every rank (test / IO helper, not on the step path)."""
        u, v = (t.contiguous() for t in self.local())
        if self.g.world == 1:
            return u.clone(), v.clone()
        outs = []
        for t in (u, v):
            parts = [torch.empty_like(t) for _ in range(self.g.world)]
            dist.all_gather(parts, t, group=self.comm.group)
            outs.append(torch.cat(parts, dim=1))
        return outs[0], outs[1]

    # ------------------------------------------------------------------ the step
    def _project_tail(self, out, p=None):
        c, be = self.comm, self.be
        c.spectral_exchange(self.send, self.recv, 0)
        be.stage_b(self.recv, self.send2)
        c.spectral_exchange(self.send2, self.recv2, 1)
        be.stage_c(self.recv2, out, p)
        self.exchanges += 2

    def project(self, want_p: bool = False):
        """P(state) in place (setup helper: initial conditions, tests)."""
        self.comm.halo_exchange(self.state, 1, 0, True)
        self.exchanges += 1
        self.be.divergence_rfft(self.state, self.send)
        p = self.be.new_pressure() if want_p else None
        self._project_tail(self.state, p)
        return p

    def step(self, nsteps: int = 1, apply_cnn: bool = False, input_scale: float = 1.0, want_p: bool = False):
        be, c = self.be, self.comm
        srcs = (self.state, self.w1, self.w2, self.w1)
        outs = (self.w1, self.w2, self.w1, self.state)
        for it in range(nsteps):
            for s in range(4):
                c.halo_exchange(srcs[s], STAGE_HALO_LO, STAGE_HALO_HI, True)
                self.exchanges += 1
                be.stage_a(s, srcs[s], self.state, self.acc, outs[s], self.send)
                p = None
                if s == 3 and want_p and it == nsteps - 1:
                    p = self.pressure = be.new_pressure()
                self._project_tail(outs[s], p)
            if apply_cnn:
                c.halo_exchange(self.state, CNN_HALO, CNN_HALO, False)
                self.exchanges += 1
                be.cnn_correct(self.state, input_scale)
                c.fence()            # p2p: the next step's halo pushes must not land while a neighbour's network still runs
