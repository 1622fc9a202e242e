"""NVLink peer-mapped ("symmetric") workspaces.

Every rank of a process group allocates the same-sized device buffer and maps every peer's buffer, so kernels can
``ld/st/red.global`` (and ``cp.async.bulk``) directly on peer memory through NVSwitch.  The first ``SIGNAL_BYTES`` of each
buffer are 32-bit flags / counters used by the in-kernel protocols (system-scope release/acquire); the rest is data.

Two substrates behind one interface:

* **vmm** (default when the driver allows it): ``cuMemCreate`` allocations exported as POSIX file descriptors (sent to
  the peers over AF_UNIX sockets, SCM_RIGHTS) plus an NVSwitch **multicast object** (``cuMulticastCreate`` /
  ``cuMulticastBindMem``) bound to every rank's buffer: ``mc_ptr`` addresses all replicas at once —
  ``multimem.ld_reduce`` sums them inside the switch, ``multimem.st`` writes them all (NVLS).  ``csrc/symm_vmm.cu``.
* **ipc**: ``cudaMalloc`` + ``cudaIpcGetMemHandle`` (no multicast); the fallback when VMM / multicast set-up fails.

This is the substrate for the fused GEMM+collective kernels (``ops/comm.py``) and the fused data-parallel / ZeRO-1
kernels; NCCL stays the bootstrap and the fallback.
"""
from __future__ import annotations

import os
import socket
import threading
import uuid
from typing import List, Optional

import torch
import torch.distributed as dist

SIGNAL_BYTES = 64 * 1024
# u32 indices inside the signal area
SIG_BARRIER = 0        # [0,16): two barrier phases x 8 peers; [16,18): grid counters (see csrc/comm.cu)
SIG_AG_READY = 64      # [64,72): "rank src's shard for all-gather epoch e is readable"
SIG_RS_ARRIVE = 128    # [128,136): tiles pushed into my reduce-scatter staging by rank src
SIG_CHUNK_CTR = 192    # [192,200): local: comm CTAs that finished copying chunk c
SIG_AUX = 256          # [256,264): small-collective counters (embedding scatter, statistics exchange)


_NODE_CACHE = {}


def peers_share_a_node(parallel_context, parallel_mode) -> bool:
    """True when every rank of the group runs on this host and the group fits one NVSwitch domain — the precondition of
    peer mapping.  Groups that span hosts (e.g. the DATA group of a multi-node job) keep the NCCL paths.
    Collective over the group (one ``all_gather_object`` of the host identities, cached per group)."""
    from pipegoose_b200.constants import MAX_NVLINK_PEERS

    ranks = tuple(parallel_context.get_ranks_in_group(parallel_mode))
    if ranks in _NODE_CACHE:
        return _NODE_CACHE[ranks]
    if len(ranks) == 1:
        same = True
    else:
        boot = ""
        try:
            boot = open("/proc/sys/kernel/random/boot_id").read().strip()
        except OSError:
            pass
        me = (socket.gethostname(), boot, os.environ.get("PIPEGOOSE_B200_FAKE_NODE", ""))
        everyone = [None] * len(ranks)
        dist.all_gather_object(everyone, me, group=parallel_context.get_group(parallel_mode))
        same = all(e == everyone[0] for e in everyone) and len(ranks) <= MAX_NVLINK_PEERS
    _NODE_CACHE[ranks] = same
    return same


# ----------------------------------------------------------------------------------------------
# file descriptors between the ranks of a node (AF_UNIX + SCM_RIGHTS, abstract socket names)
# ----------------------------------------------------------------------------------------------
def exchange_fds(group, rank: int, world: int, my_fds: List[int], timeout_s: float = 60.0) -> List[List[int]]:
    """Every rank offers ``my_fds`` (same count on all ranks; ``-1`` entries are skipped and come back as ``-1``);
    returns ``fds[r]`` = rank r's descriptors, duplicated into this process (``fds[rank]`` is ``my_fds``).
    Collective over ``group``."""
    name = f"\0pgb200-{uuid.uuid4().hex}"
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.bind(name)
    server.listen(world)
    server.settimeout(timeout_s)
    names = [None] * world
    dist.all_gather_object(names, name, group=group)
    valid = [f for f in my_fds if f >= 0]
    errors = []

    def serve():
        try:
            for _ in range(world - 1):
                conn, _addr = server.accept()
                with conn:
                    conn.settimeout(timeout_s)
                    if valid:
                        socket.send_fds(conn, [b"f"], valid)
                    else:
                        conn.sendall(b"n")
        except Exception as e:  # pragma: no cover - reported by the caller
            errors.append(e)

    t = threading.Thread(target=serve, daemon=True)
    t.start()
    out: List[List[int]] = [[] for _ in range(world)]
    out[rank] = list(my_fds)
    try:
        for r in range(world):
            if r == rank:
                continue
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.settimeout(timeout_s)
                c.connect(names[r])
                _msg, fds, _flags, _addr = socket.recv_fds(c, 16, len(my_fds))
                out[r] = list(fds)
    except Exception as e:
        errors.append(e)
    t.join(timeout_s)
    server.close()
    if errors:
        raise errors[0]
    # (no collective here: a rank that failed above must not leave its peers in a barrier it never reaches — the caller
    #  agrees on the outcome with one all-gather)
    return out


class SymmetricWorkspace:
    def __init__(self, parallel_context, parallel_mode, nbytes: int):
        from pipegoose_b200.ops import native

        self._n = native()
        self.ctx = parallel_context
        self.mode = parallel_mode
        self.group = parallel_context.get_group(parallel_mode)
        self.world = parallel_context.get_world_size(parallel_mode)
        self.rank = parallel_context.get_local_rank(parallel_mode)
        self.nbytes = int(nbytes)
        self.device_index = torch.cuda.current_device()
        self.mc_ptr = 0            # multicast address of the whole buffer (0: no NVLS)
        self.backend = "ipc"
        self._closed = False
        self._vmm = None
        want = os.environ.get("PIPEGOOSE_B200_SYMM", "auto")   # auto / vmm / ipc
        ok = False
        if want != "ipc" and self.world > 1:
            ok = self._init_vmm(SIGNAL_BYTES + ((self.nbytes + 1023) // 1024) * 1024)
            if not ok and want == "vmm":
                raise RuntimeError("PIPEGOOSE_B200_SYMM=vmm: VMM / multicast symmetric memory could not be set up")
        if not ok:
            self._init_ipc(SIGNAL_BYTES + ((self.nbytes + 1023) // 1024) * 1024)
        self._bytes = self._n.tensor_from_ptr(self._local_ptr, self._total, self.device_index)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    # ------------------------------------------------------------------ substrates
    def _init_ipc(self, total: int):
        self._total = total
        self._local_ptr, handle = self._n.symm_alloc(total)
        handles: List[bytes] = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        self.peer_ptrs: List[int] = []
        for r, h in enumerate(handles):
            self.peer_ptrs.append(self._local_ptr if r == self.rank else self._n.symm_open(h))
        self.backend = "ipc"

    def _agree(self, ok: bool) -> bool:
        """All ranks take the same branch: a step counts as done only if it succeeded everywhere."""
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=self.group)
        return all(flags)

    def _init_vmm(self, total: int) -> bool:
        n = self._n
        try:
            usable, mc_ok, gran = n.vmm_probe(self.world)
        except Exception:
            usable, mc_ok, gran = False, False, 0
        if not self._agree(usable and gran > 0):
            return False
        grans = [None] * self.world
        dist.all_gather_object(grans, int(gran), group=self.group)
        gran = max(grans)
        total = (total + gran - 1) // gran * gran
        state = {"maps": [], "mc_handle": 0, "mc_map": 0}
        try:
            ptr, fd, handle = n.vmm_alloc(total)
            state["maps"].append((ptr, handle))
            ok = True
        except Exception as e:
            print(f"[pipegoose_b200] vmm_alloc failed: {e}", flush=True)
            ptr, fd, handle, ok = 0, -1, 0, False
        if not self._agree(ok):
            self._release_vmm(state, total)
            return False
        mc_fd, mc_handle = -1, 0
        want_mc = mc_ok and os.environ.get("PIPEGOOSE_B200_NVLS", "1") == "1"
        mc_flags = [None] * self.world
        dist.all_gather_object(mc_flags, bool(want_mc), group=self.group)
        want_mc = all(mc_flags)
        if want_mc and self.rank == 0:
            try:
                mc_fd, mc_handle = n.mc_create(self.world, total)
            except Exception as e:
                print(f"[pipegoose_b200] multicast object not created ({e}): peer mappings without NVLS", flush=True)
                mc_fd, mc_handle = -1, 0
        try:
            fds = exchange_fds(self.group, self.rank, self.world, [fd, mc_fd if self.rank == 0 else -1])
            ok = True
        except Exception as e:
            print(f"[pipegoose_b200] file-descriptor exchange failed: {e}", flush=True)
            fds, ok = None, False
        if fd >= 0:
            os.close(fd)
        if mc_fd >= 0:
            os.close(mc_fd)
        if not self._agree(ok):
            self._release_vmm(state, total)
            return False
        peer_ptrs = [0] * self.world
        peer_ptrs[self.rank] = ptr
        ok = True
        try:
            for r in range(self.world):
                if r == self.rank:
                    continue
                p, h = n.vmm_import(fds[r][0], total)     # (closes the descriptor)
                state["maps"].append((p, h))
                peer_ptrs[r] = p
        except Exception as e:
            print(f"[pipegoose_b200] vmm_import failed: {e}", flush=True)
            ok = False
        if not self._agree(ok):
            self._release_vmm(state, total)
            return False
        # multicast object: rank 0 created it; everybody imports, adds its device, then binds its buffer
        have_mc = [None] * self.world
        dist.all_gather_object(have_mc, bool(want_mc and (mc_handle != 0 if self.rank == 0 else len(fds[0]) > 1)),
                               group=self.group)
        mc_ptr = 0
        if all(have_mc):
            ok = True
            try:
                if self.rank != 0:
                    mc_handle = n.mc_import(fds[0][1])
                n.mc_add_device(mc_handle)
            except Exception as e:
                print(f"[pipegoose_b200] multicast add_device failed: {e}", flush=True)
                ok = False
            if self._agree(ok):      # every device was added before anyone binds
                try:
                    mc_ptr = n.mc_bind(mc_handle, handle, total)
                    state["mc_map"] = mc_ptr
                except Exception as e:
                    print(f"[pipegoose_b200] multicast bind failed: {e}", flush=True)
                    ok = False
                if not self._agree(ok):
                    mc_ptr = 0
            state["mc_handle"] = mc_handle
        self._total = total
        self._local_ptr = ptr
        self.peer_ptrs = peer_ptrs
        self.mc_ptr = mc_ptr
        self._vmm = state
        self.backend = "vmm+nvls" if mc_ptr else "vmm"
        return True

    def _release_vmm(self, state, total):
        for p, h in state.get("maps", []):
            try:
                self._n.vmm_unmap(p, total, h)
            except Exception:
                pass
        state["maps"] = []

    # ------------------------------------------------------------------ addressing
    def data_ptr(self, rank: int, offset: int = 0) -> int:
        return self.peer_ptrs[rank] + SIGNAL_BYTES + offset

    def mc_data_ptr(self, offset: int = 0) -> int:
        """Multicast address of data offset ``offset`` (0 when the workspace has no multicast object)."""
        return self.mc_ptr + SIGNAL_BYTES + offset if self.mc_ptr else 0

    def sig_ptr(self, rank: int, index: int) -> int:
        return self.peer_ptrs[rank] + 4 * index

    def local_tensor(self, offset: int, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        assert offset % 16 == 0 and offset + nbytes <= self.nbytes, "symmetric workspace overflow"
        start = SIGNAL_BYTES + offset
        return self._bytes[start:start + nbytes].view(dtype).view(*shape)

    def signals(self) -> torch.Tensor:
        return self._bytes[:SIGNAL_BYTES].view(torch.int32)

    def close(self):
        if self._closed:
            return
        self._closed = True
        try:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
        except Exception:
            pass
        self._bytes = None
        if self._vmm is not None:
            if self._vmm.get("mc_map"):
                try:
                    self._n.vmm_unmap(self._vmm["mc_map"], self._total, self._vmm.get("mc_handle", 0))
                except Exception:
                    pass
            self._release_vmm(self._vmm, self._total)
            return
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                self._n.symm_close(p)
        self._n.symm_free(self._local_ptr)
