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
          (fused sweep + exchange); a per-step barrier orders the buffer reuse.
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

    def _step_p2p(self):
        import torch.distributed as dist
        # the sweep stores the leaving populations straight into the neighbours' ghost planes of the
        # buffer being written; the barrier orders "my writes landed / your reads finished"
        self.dom.step(1)
        self.stream.synchronize()
        dist.barrier(group=self.group)

    # ---- public ----------------------------------------------------------------------
    def step(self, n=1):
        if self.transport == "none":
            self.dom.step(n)
            return
        fn = self._step_nccl if self.transport == "nccl" else self._step_p2p
        with self.torch.cuda.stream(self.stream):
            for _ in range(n):
                fn()

    def close(self):
        self.dom.close()
