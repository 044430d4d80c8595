"""NUMA placement of the pinned host buffers the end-to-end path copies from (HostFedPipeline.pin_host_batch).

On an 8-GPU node the GPUs hang off two CPU sockets.  A pinned buffer lands on the NUMA node of the thread that first touches
it; when every rank runs on the same socket (a container whose cpuset is one socket), half of the GPUs read their step's
9.8 MB across the socket interconnect, and all eight copies share one socket's memory controllers.  ``numa_preferred(node)``
sets the calling thread's memory policy (``set_mempolicy(MPOL_PREFERRED)``: a preference, never a failure) around the
allocation, ``gpu_numa_node`` reads the node a GPU is attached to from sysfs.  Host-side plumbing only; nothing here touches
the device path.  Linux only; every function degrades to a no-op when the syscall or the sysfs entry is missing."""
import contextlib
import ctypes
import os

MPOL_DEFAULT, MPOL_PREFERRED = 0, 1
_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237}.get(os.uname().machine)
_SYS_MOVE_PAGES = {"x86_64": 279, "aarch64": 239}.get(os.uname().machine)
_libc = ctypes.CDLL(None, use_errno=True)


def gpu_numa_node(device_index):
    """NUMA node of CUDA device ``device_index`` (-1 when unknown)."""
    bdf = None
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
    except Exception:
        bdf = None
    if bdf is None:
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        except Exception:
            return -1
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:      # NVML prints an 8-digit domain, sysfs a 4-digit one
        bdf = bdf[4:]
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def mems_allowed():
    try:
        with open("/proc/self/status") as f:
            for line in f:
                if line.startswith("Mems_allowed_list"):
                    return line.split(":")[1].strip()
    except Exception:
        pass
    return None


def _node_allowed(node):
    allowed = mems_allowed()
    if allowed is None:
        return False
    for part in allowed.split(","):
        lo, _, hi = part.partition("-")
        if lo.isdigit() and int(lo) <= node <= int(hi or lo):
            return True
    return False


@contextlib.contextmanager
def numa_preferred(node):
    """Allocations first touched inside the block prefer NUMA node ``node`` (no-op for node < 0 / unsupported / not allowed)."""
    ok = False
    if node is not None and node >= 0 and _SYS_SET_MEMPOLICY is not None and _node_allowed(node):
        mask = ctypes.c_ulong(1 << node)
        ok = _libc.syscall(_SYS_SET_MEMPOLICY, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask))) == 0
    try:
        yield ok
    finally:
        if ok:
            _libc.syscall(_SYS_SET_MEMPOLICY, MPOL_DEFAULT, None, ctypes.c_ulong(0))


def node_of_buffer(tensor):
    """NUMA node of the first page of a host tensor (move_pages query; -1 when unknown)."""
    if _SYS_MOVE_PAGES is None:
        return -1
    page = ctypes.c_void_p(tensor.data_ptr() & ~4095)
    status = ctypes.c_int(-1)
    rc = _libc.syscall(_SYS_MOVE_PAGES, 0, ctypes.c_ulong(1), ctypes.byref(page), None, ctypes.byref(status), 0)
    return int(status.value) if rc == 0 else -1
