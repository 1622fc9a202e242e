// Symmetric memory on the CUDA virtual-memory-management API, with an NVSwitch multicast object bound to it.
//
//   pg_vmm_alloc      cuMemCreate (POSIX-fd shareable) + reserve + map + access  -> local pointer, fd to send to peers
//   pg_vmm_import     cuMemImportFromShareableHandle(fd) + reserve + map + access -> peer pointer (NVLink ld/st/red/TMA)
//   pg_mc_create      rank 0: cuMulticastCreate(world devices, size)              -> fd to send to peers
//   pg_mc_import      other ranks: import the multicast object from the fd
//   pg_mc_add_device  every rank: cuMulticastAddDevice (all ranks must have done it before anyone binds)
//   pg_mc_bind        every rank: cuMulticastBindMem(local allocation) + map the multicast object -> multicast pointer:
//                     multimem.ld_reduce on it sums the replicas inside the switch, multimem.st / multimem.red write
//                     all of them (NVLS)
//
// The file descriptors travel between the ranks of a node over AF_UNIX sockets (SCM_RIGHTS; distributed/symmetric.py).
// Driver entry points are resolved at run time (cudaGetDriverEntryPoint), so the extension links without libcuda.
// Every function returns 0 on success; on failure it prints the driver call that failed and returns -1 — the Python
// side then falls back to the cudaIpc workspace (no multicast).
#include "launch.h"

#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <unistd.h>

namespace {

template <typename F>
F drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) {
    fprintf(stderr, "pipegoose_b200: driver entry point %s not available\n", name);
    return nullptr;
  }
  return reinterpret_cast<F>(p);
}

const char* err_str(CUresult r) {
  typedef CUresult (*Fn)(CUresult, const char**);
  static Fn fn = drv<Fn>("cuGetErrorString");
  const char* s = nullptr;
  if (fn != nullptr && fn(r, &s) == CUDA_SUCCESS && s != nullptr) return s;
  return "unknown";
}

#define PG_DRV(call, name)                                                                   \
  do {                                                                                       \
    CUresult r__ = (call);                                                                   \
    if (r__ != CUDA_SUCCESS) {                                                               \
      fprintf(stderr, "pipegoose_b200: %s failed: %d (%s)\n", name, (int)r__, err_str(r__)); \
      return -1;                                                                             \
    }                                                                                        \
  } while (0)

typedef CUresult (*PFN_cuMemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*PFN_cuMemRelease)(CUmemGenericAllocationHandle);
typedef CUresult (*PFN_cuMemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*PFN_cuMemAddressFree)(CUdeviceptr, size_t);
typedef CUresult (*PFN_cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*PFN_cuMemUnmap)(CUdeviceptr, size_t);
typedef CUresult (*PFN_cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
typedef CUresult (*PFN_cuMemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
typedef CUresult (*PFN_cuMemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
typedef CUresult (*PFN_cuMemGetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
typedef CUresult (*PFN_cuMulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
typedef CUresult (*PFN_cuMulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
typedef CUresult (*PFN_cuMulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                           unsigned long long);
typedef CUresult (*PFN_cuMulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
typedef CUresult (*PFN_cuDeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);

int cur_device() {
  int dev = 0;
  cudaFree(nullptr);  // primary context bound to this thread
  cudaGetDevice(&dev);
  return dev;
}

CUmemAllocationProp mem_prop(int dev) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int map_handle(CUmemGenericAllocationHandle h, size_t size, int dev, void** ptr) {
  static auto reserve = drv<PFN_cuMemAddressReserve>("cuMemAddressReserve");
  static auto mapfn = drv<PFN_cuMemMap>("cuMemMap");
  static auto access = drv<PFN_cuMemSetAccess>("cuMemSetAccess");
  if (!reserve || !mapfn || !access) return -1;
  CUdeviceptr va = 0;
  PG_DRV(reserve(&va, size, 0, 0, 0), "cuMemAddressReserve");
  PG_DRV(mapfn(va, size, 0, h, 0), "cuMemMap");
  CUmemAccessDesc desc;
  memset(&desc, 0, sizeof(desc));
  desc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  desc.location.id = dev;
  desc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  PG_DRV(access(va, size, &desc, 1), "cuMemSetAccess");
  *ptr = reinterpret_cast<void*>(va);
  return 0;
}

}  // namespace

// mc_supported: the device can join NVSwitch multicast objects; gran: allocation granularity to round sizes to
// (covers the multicast minimum granularity when multicast is supported)
extern "C" int pg_vmm_probe(int world, int* mc_supported, int64_t* gran) {
  const int dev = cur_device();
  static auto get_attr = drv<PFN_cuDeviceGetAttribute>("cuDeviceGetAttribute");
  static auto mem_gran = drv<PFN_cuMemGetGranularity>("cuMemGetAllocationGranularity");
  if (!get_attr || !mem_gran) return -1;
  int vmm = 0, fd_ok = 0, mc = 0;
  PG_DRV(get_attr(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev), "cuDeviceGetAttribute(VMM)");
  PG_DRV(get_attr(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev), "cuDeviceGetAttribute(FD)");
  if (!vmm || !fd_ok) {
    fprintf(stderr, "pipegoose_b200: VMM %d / POSIX-fd handles %d not supported on device %d\n", vmm, fd_ok, dev);
    return -1;
  }
  if (get_attr(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) mc = 0;
  CUmemAllocationProp prop = mem_prop(dev);
  size_t g = 0;
  PG_DRV(mem_gran(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (mc) {
    static auto mc_gran = drv<PFN_cuMulticastGetGranularity>("cuMulticastGetGranularity");
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = world > 0 ? world : 1;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (mc_gran == nullptr || mc_gran(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS) {
      mc = 0;
    } else if (mg > g) {
      g = (mg + g - 1) / g * g;
    }
  }
  *mc_supported = mc;
  *gran = static_cast<int64_t>(g);
  return 0;
}

// nbytes must be a multiple of the granularity pg_vmm_probe reported.  The memory is zero-filled.
extern "C" int pg_vmm_alloc(int64_t nbytes, void** ptr, int* fd, uint64_t* handle) {
  const int dev = cur_device();
  static auto create = drv<PFN_cuMemCreate>("cuMemCreate");
  static auto exportfn = drv<PFN_cuMemExport>("cuMemExportToShareableHandle");
  if (!create || !exportfn) return -1;
  CUmemAllocationProp prop = mem_prop(dev);
  CUmemGenericAllocationHandle h;
  PG_DRV(create(&h, static_cast<size_t>(nbytes), &prop, 0), "cuMemCreate");
  if (map_handle(h, static_cast<size_t>(nbytes), dev, ptr) != 0) return -1;
  int f = -1;
  PG_DRV(exportfn(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  if (cudaMemset(*ptr, 0, static_cast<size_t>(nbytes)) != cudaSuccess) return -1;
  *fd = f;
  *handle = static_cast<uint64_t>(h);
  return 0;
}

// map a peer's allocation (received as a file descriptor; the descriptor is closed here)
extern "C" int pg_vmm_import(int fd, int64_t nbytes, void** ptr, uint64_t* handle) {
  const int dev = cur_device();
  static auto importfn = drv<PFN_cuMemImport>("cuMemImportFromShareableHandle");
  if (!importfn) return -1;
  CUmemGenericAllocationHandle h;
  PG_DRV(importfn(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
         "cuMemImportFromShareableHandle");
  close(fd);
  if (map_handle(h, static_cast<size_t>(nbytes), dev, ptr) != 0) return -1;
  *handle = static_cast<uint64_t>(h);
  return 0;
}

extern "C" int pg_vmm_unmap(void* ptr, int64_t nbytes, uint64_t handle) {
  static auto unmap = drv<PFN_cuMemUnmap>("cuMemUnmap");
  static auto freeva = drv<PFN_cuMemAddressFree>("cuMemAddressFree");
  static auto release = drv<PFN_cuMemRelease>("cuMemRelease");
  if (!unmap || !freeva || !release) return -1;
  if (ptr != nullptr) {
    unmap(reinterpret_cast<CUdeviceptr>(ptr), static_cast<size_t>(nbytes));
    freeva(reinterpret_cast<CUdeviceptr>(ptr), static_cast<size_t>(nbytes));
  }
  if (handle != 0) release(static_cast<CUmemGenericAllocationHandle>(handle));
  return 0;
}

extern "C" int pg_mc_create(int world, int64_t nbytes, int* fd, uint64_t* mc_handle) {
  cur_device();
  static auto create = drv<PFN_cuMulticastCreate>("cuMulticastCreate");
  static auto exportfn = drv<PFN_cuMemExport>("cuMemExportToShareableHandle");
  if (!create || !exportfn) return -1;
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = world;
  mp.size = static_cast<size_t>(nbytes);
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle h;
  PG_DRV(create(&h, &mp), "cuMulticastCreate");
  int f = -1;
  PG_DRV(exportfn(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle(multicast)");
  *fd = f;
  *mc_handle = static_cast<uint64_t>(h);
  return 0;
}

extern "C" int pg_mc_import(int fd, uint64_t* mc_handle) {
  cur_device();
  static auto importfn = drv<PFN_cuMemImport>("cuMemImportFromShareableHandle");
  if (!importfn) return -1;
  CUmemGenericAllocationHandle h;
  PG_DRV(importfn(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
         "cuMemImportFromShareableHandle(multicast)");
  close(fd);
  *mc_handle = static_cast<uint64_t>(h);
  return 0;
}

extern "C" int pg_mc_add_device(uint64_t mc_handle) {
  const int dev = cur_device();
  static auto add = drv<PFN_cuMulticastAddDevice>("cuMulticastAddDevice");
  if (!add) return -1;
  PG_DRV(add(static_cast<CUmemGenericAllocationHandle>(mc_handle), static_cast<CUdevice>(dev)), "cuMulticastAddDevice");
  return 0;
}

// bind this rank's allocation at offset 0 of the multicast object and map the object; every rank of the group must
// have called pg_mc_add_device before the first bind
extern "C" int pg_mc_bind(uint64_t mc_handle, uint64_t mem_handle, int64_t nbytes, void** mc_ptr) {
  const int dev = cur_device();
  static auto bind = drv<PFN_cuMulticastBindMem>("cuMulticastBindMem");
  if (!bind) return -1;
  PG_DRV(bind(static_cast<CUmemGenericAllocationHandle>(mc_handle), 0, static_cast<CUmemGenericAllocationHandle>(mem_handle),
              0, static_cast<size_t>(nbytes), 0),
         "cuMulticastBindMem");
  return map_handle(static_cast<CUmemGenericAllocationHandle>(mc_handle), static_cast<size_t>(nbytes), dev, mc_ptr);
}
