"""z-slab decomposition across GPUs, one process per GPU (torch.distributed for the plumbing).

Replaces the intent of the reference's parallel.h:11-23 / Domain::create_subdomain
(domain.hpp:197-248), which never exchanged anything.  z is the slowest index
(domain.hpp:61-64), so in the f[q][z][y][x] layout each population of an x-y plane
is one contiguous block.  After each sweep, the c_z=+1 populations of a slab's top
interior plane must reach the upper neighbour's bottom ghost plane and the c_z=-1
populations of its bottom interior plane the lower neighbour's top ghost plane:
5 (D3Q15/19) or 9 (D3Q27) planes per direction per interface.

Transports
  "nccl"  split-phase: edge planes are swept first, their halo planes go out with
          grouped ncclSend/ncclRecv (torch.distributed.batch_isend_irecv) while the
          interior planes are swept on the compute stream.
  "p2p"   the sweep kernel itself stores the leaving populations into the
          neighbour's ghost planes through CUDA-IPC peer mappings over NVLink
          (fused sweep + exchange); per-side arrival counters in device memory,
          bumped with system-scope release stores and awaited by a one-thread
          kernel in front of the next sweep, order the buffer reuse -- no host
          synchronisation inside the time loop.
"""
import os

from . import capi

DOWN, UP = capi.DOWN, capi.UP


def partition(zl_global, world):
    """Balanced contiguous z ranges: [(z_first, zl_local)] with 1-based z_first."""
    if world < 1 or zl_global < world:
        raise ValueError("cannot split %d planes over %d slabs" % (zl_global, world))
    base, extra = divmod(zl_global, world)
    out, z = [], 1
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((z, n))
        z += n
    return out


def neighbours(rank, world, periodic_z=False):
    """(down, up) ranks or None."""
    down = rank - 1 if rank > 0 else (world - 1 if periodic_z and world > 1 else None)
    up = rank + 1 if rank < world - 1 else (0 if periodic_z and world > 1 else None)
    return down, up


def halo_ops(dist, send, recv, down, up):
    """Build the P2POp list of one exchange.

    send/recv: {DOWN: [tensors], UP: [tensors]}.  Planes sent UP by rank r are received as the DOWN
    planes of rank r+1, in the same order (k-th c_z=+1 population), and vice versa.
    """
    ops = []
    # fixed global order (all "up-going" traffic first) so that paired ranks post matching sequences
    if up is not None:
        ops += [dist.P2POp(dist.isend, t, up) for t in send[UP]]
    if down is not None:
        ops += [dist.P2POp(dist.irecv, t, down) for t in recv[DOWN]]
    if down is not None:
        ops += [dist.P2POp(dist.isend, t, down) for t in send[DOWN]]
    if up is not None:
        ops += [dist.P2POp(dist.irecv, t, up) for t in recv[UP]]
    return ops


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can wrap library-owned device memory."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2,
                                         "strides": None}


class SlabRunner:
    """One rank's slab of a global Domain plus its halo exchange."""

    def __init__(self, Q, xl, yl, zl_global, tau, boxes, rank=0, world=1, device=None, transport="nccl",
                 periodic_z=False, exact=False, group=None):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.z_first, self.zl = partition(zl_global, world)[rank]
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", 0))
        self.device = device
        torch.cuda.set_device(device)
        self.dom = capi.Domain(Q, xl, yl, zl_global, tau, device=device, z_first=self.z_first, zl_local=self.zl,
                               exact=exact)
        if boxes:
            self.dom.set_boxes(boxes)
        self.down, self.up = neighbours(rank, world, periodic_z)
        self.transport = transport if world > 1 else "none"
        self.group = group
        # a dedicated non-default stream: the library's kernels, the NCCL ordering events and the
        # timing events of bench.py all refer to it
        self.stream = torch.cuda.Stream(device=device)
        self.dom.set_stream(self.stream.cuda_stream)
        if self.transport == "nccl":
            self._wrap_planes()
        elif self.transport == "p2p":
            self._connect_peers()

    # ---- nccl transport ------------------------------------------------------------
    def _wrap_planes(self):
        torch = self.torch
        n_q, plane_bytes = self.dom.halo_layout()
        n = plane_bytes // 8
        self.planes = {}
        self._keep = []
        for buf in (0, 1):
            for recv in (False, True):
                d = {}
                for side in (DOWN, UP):
                    ts = []
                    for k in range(n_q):
                        arr = _DevArray(self.dom.halo_plane(buf, side, k, recv), n)
                        self._keep.append(arr)
                        ts.append(torch.as_tensor(arr, device="cuda:%d" % self.device))
                    d[side] = ts
                self.planes[(buf, recv)] = d

    def _step_nccl(self):
        import torch.distributed as dist
        dom = self.dom
        dom.step_edges()
        buf = dom.dst_buffer()
        ops = halo_ops(dist, self.planes[(buf, False)], self.planes[(buf, True)], self.down, self.up)
        reqs = dist.batch_isend_irecv(ops) if ops else []
        dom.step_interior()
        for r in reqs:
            r.wait()
        dom.step_finish()

    # ---- p2p transport -------------------------------------------------------------
    def _connect_peers(self):
        import torch.distributed as dist
        blobs = [None] * self.world
        dist.all_gather_object(blobs, self.dom.export(), group=self.group)
        if self.up is not None:
            self.dom.connect(UP, blobs[self.up])
        if self.down is not None:
            self.dom.connect(DOWN, blobs[self.down])
        dist.barrier(group=self.group)

    def _step_p2p(self, n):
        # the sweep stores the leaving populations straight into the neighbours' ghost planes of the
        # buffer being written; the library's device-side hand-shake orders the buffer reuse
        self.dom.step(n)

    # ---- public ----------------------------------------------------------------------
    def step(self, n=1):
        if self.transport == "none" or n == 0:
            self.dom.step(n)
            return
        if self.transport == "p2p":
            self._step_p2p(n)
            return
        with self.torch.cuda.stream(self.stream):
            for _ in range(n):
                self._step_nccl()

    def sync(self):
        self.dom.sync()

    def prepare_readback(self):
        """full edge planes (all Q populations) across the cuts before download()/macroscopic(), so that boundary
        cells next to a cut are materialised exactly like on one GPU"""
        import torch.distributed as dist
        self.dom.sync()
        if self.transport == "p2p":
            dist.barrier(group=self.group)
            self.dom.halo_push_all()
            self.dom.sync()
            dist.barrier(group=self.group)
            self.dom.halo_pushed()
        elif self.transport == "nccl":
            torch = self.torch
            n = self.dom.halo_layout()[1] // 8
            keep, ops = [], []

            def plane(side, q, recv):
                arr = _DevArray(self.dom.edge_plane(side, q, recv), n)
                keep.append(arr)
                return torch.as_tensor(arr, device="cuda:%d" % self.device)
            for q in range(self.dom.Q):
                if self.up is not None:
                    ops.append(dist.P2POp(dist.isend, plane(UP, q, False), self.up))
                    ops.append(dist.P2POp(dist.irecv, plane(UP, q, True), self.up))
                if self.down is not None:
                    ops.append(dist.P2POp(dist.isend, plane(DOWN, q, False), self.down))
                    ops.append(dist.P2POp(dist.irecv, plane(DOWN, q, True), self.down))
            with torch.cuda.stream(self.stream):
                for r in (dist.batch_isend_irecv(ops) if ops else []):
                    r.wait()
            self.dom.sync()
            self.dom.halo_pushed()

    def close(self):
        """collective for the p2p transport: nobody frees its slab while a neighbour still maps it"""
        if self.transport == "p2p":
            import torch.distributed as dist
            self.dom.sync()
            dist.barrier(group=self.group)
            self.dom.disconnect()
            dist.barrier(group=self.group)
        self.dom.close()


class LocalSlabStack:
    """N slabs driven from ONE process (what the C++ Domain does for `gpus = N`): slabs may sit on
    different GPUs (peer access over NVLink) or, for testing the exchange logic on a single GPU, on
    the same device.  axis = capi.AXIS_Z splits into z-slabs (x-y planes), capi.AXIS_Y into y-slabs
    (x-z planes) -- for domains whose z extent is shorter than the number of slabs, or flat ones."""

    def __init__(self, Q, xl, yl, zl, tau, boxes, n_slabs, devices=None, exact=False, fluid_mask=None,
                 periodic_z=False, axis=capi.AXIS_Z):
        import numpy as np
        self.np = np
        self.Q, self.xl, self.yl, self.zl, self.axis = Q, xl, yl, zl, axis
        self.periodic_z = periodic_z and n_slabs > 1      # periodic along the split axis: a ring of slabs
        self.ranges = partition(yl if axis == capi.AXIS_Y else zl, n_slabs)
        devices = devices or [0] * n_slabs
        self.slabs = [capi.Domain(Q, xl, yl, zl, tau, device=devices[r], z_first=zf, zl_local=nz, exact=exact, axis=axis)
                      for r, (zf, nz) in enumerate(self.ranges)]
        for (zf, nz), s in zip(self.ranges, self.slabs):
            if fluid_mask is not None:      # every slab also needs the rows of its neighbours' edge planes
                s.set_fluid_mask_global(fluid_mask)
            if boxes:
                s.set_boxes(boxes)
        for r in range(n_slabs):
            down, up = neighbours(r, n_slabs, periodic_z)
            if up is not None:
                self.slabs[r].connect_local(UP, self.slabs[up])
            if down is not None:
                self.slabs[r].connect_local(DOWN, self.slabs[down])

    def _cut(self, a, lo, hi):
        """planes lo..hi-1 of the split axis of an array shaped [z, y, ...]"""
        return a[:, lo:hi] if self.axis == capi.AXIS_Y else a[lo:hi]

    def upload(self, f):
        """f: [(xl+2)(yl+2)(zl+2), Q] global AoS; every slab gets its planes plus both ghost planes
        (for a periodic ring the two global ghost planes are first filled with their wrapped images)"""
        np = self.np
        f = np.array(f, dtype=np.float64).reshape(self.zl + 2, self.yl + 2, (self.xl + 2) * self.Q)
        n = self.yl if self.axis == capi.AXIS_Y else self.zl
        if self.periodic_z:
            self._cut(f, 0, 1)[...] = self._cut(f, n, n + 1)
            self._cut(f, n + 1, n + 2)[...] = self._cut(f, 1, 2)
        for (zf, nz), s in zip(self.ranges, self.slabs):
            s.upload(np.ascontiguousarray(self._cut(f, zf - 1, zf + nz + 1)))

    def step(self, n=1):
        capi.step_group(self.slabs, n)

    def sync(self):
        for s in self.slabs:
            s.sync()

    def prepare_readback(self):
        """push full edge planes across the cuts so that boundary cells next to a cut read back exactly"""
        self.sync()
        for s in self.slabs:
            s.halo_push_all()
        self.sync()
        for s in self.slabs:
            s.halo_pushed()

    def download(self):
        """global AoS populations; interface ghost planes are taken from their owners"""
        np = self.np
        row = (self.xl + 2) * self.Q
        out = np.empty((self.zl + 2, self.yl + 2, row))
        self.prepare_readback()
        for i, ((zf, nz), s) in enumerate(zip(self.ranges, self.slabs)):
            loc = s.download().reshape(s.zl + 2, s.yl + 2, row)
            lo = 0 if i == 0 else 1
            hi = nz + 2 if i == len(self.slabs) - 1 else nz + 1
            self._cut(out, zf - 1 + lo, zf - 1 + hi)[...] = self._cut(loc, lo, hi)
        return out.reshape(-1, self.Q)

    def macroscopic(self):
        np = self.np
        rho = np.empty((self.zl, self.yl, self.xl))
        u = np.empty((self.zl, self.yl, self.xl, 3))
        self.prepare_readback()
        for (zf, nz), s in zip(self.ranges, self.slabs):
            r, v = s.macroscopic()
            self._cut(rho, zf - 1, zf - 1 + nz)[...] = r
            self._cut(u, zf - 1, zf - 1 + nz)[...] = v
        return rho, u

    def close(self):
        for s in self.slabs:
            s.sync()
        for s in self.slabs:
            s.disconnect()
        for s in self.slabs:
            s.close()
