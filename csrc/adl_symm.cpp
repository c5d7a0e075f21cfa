// adaptdl_b200 -- symmetric (peer-mapped) memory runtime.
//
// Owns the cross-process GPU memory mappings that the fused gradient kernels
// load/store through: physical allocations are created with the CUDA virtual
// memory management API as POSIX-fd shareable handles, the fds travel to the
// peer processes over a unix socket (SCM_RIGHTS, done by the Python side), and
// every rank maps every peer's allocation into its own address space. On
// NVSwitch systems an NVLS multicast object can be bound over the same
// physical pages for multimem.* instructions.
//
// The driver API is resolved with dlopen("libcuda.so.1") at first use so this
// library also loads on machines without a GPU driver (CPU-only CI).
//
// Elastic restarts tear all of this down with the process and rebuild it at
// the new world size (generation = ADAPTDL_NUM_RESTARTS).
#include <cuda.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <mutex>

namespace {

struct Driver {
  void* lib = nullptr;
  CUresult (*Init)(unsigned) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*DevicePrimaryCtxRetain)(CUcontext*, CUdevice) = nullptr;
  CUresult (*CtxSetCurrent)(CUcontext) = nullptr;
  CUresult (*DeviceCanAccessPeer)(int*, CUdevice, CUdevice) = nullptr;
  CUresult (*DeviceGetP2PAttribute)(int*, CUdevice_P2PAttribute, CUdevice, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                          CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle,
                                         CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*,
                                           CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle,
                     unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle,
                               size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*,
                                      CUmulticastGranularity_flags) = nullptr;
};

Driver g_drv;
std::once_flag g_once;
int g_load_status = -1;
char g_last_error[512] = "";

template <typename F>
bool sym(F& fn, const char* name) {
  fn = reinterpret_cast<F>(dlsym(g_drv.lib, name));
  return fn != nullptr;
}

void load_driver() {
  g_drv.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!g_drv.lib) g_drv.lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
  if (!g_drv.lib) {
    snprintf(g_last_error, sizeof g_last_error, "cannot load libcuda: %s", dlerror());
    g_load_status = -1;
    return;
  }
  bool ok = true;
  ok &= sym(g_drv.Init, "cuInit");
  ok &= sym(g_drv.GetErrorString, "cuGetErrorString");
  ok &= sym(g_drv.DeviceGet, "cuDeviceGet");
  ok &= sym(g_drv.DeviceGetAttribute, "cuDeviceGetAttribute");
  ok &= sym(g_drv.DevicePrimaryCtxRetain, "cuDevicePrimaryCtxRetain");
  ok &= sym(g_drv.CtxSetCurrent, "cuCtxSetCurrent");
  ok &= sym(g_drv.DeviceCanAccessPeer, "cuDeviceCanAccessPeer");
  ok &= sym(g_drv.DeviceGetP2PAttribute, "cuDeviceGetP2PAttribute");
  ok &= sym(g_drv.MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
  ok &= sym(g_drv.MemCreate, "cuMemCreate");
  ok &= sym(g_drv.MemRelease, "cuMemRelease");
  ok &= sym(g_drv.MemExportToShareableHandle, "cuMemExportToShareableHandle");
  ok &= sym(g_drv.MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
  ok &= sym(g_drv.MemAddressReserve, "cuMemAddressReserve");
  ok &= sym(g_drv.MemAddressFree, "cuMemAddressFree");
  ok &= sym(g_drv.MemMap, "cuMemMap");
  ok &= sym(g_drv.MemUnmap, "cuMemUnmap");
  ok &= sym(g_drv.MemSetAccess, "cuMemSetAccess");
  // multicast entry points are optional (driver >= 12.1)
  sym(g_drv.MulticastCreate, "cuMulticastCreate");
  sym(g_drv.MulticastAddDevice, "cuMulticastAddDevice");
  sym(g_drv.MulticastBindMem, "cuMulticastBindMem");
  sym(g_drv.MulticastGetGranularity, "cuMulticastGetGranularity");
  if (!ok) {
    snprintf(g_last_error, sizeof g_last_error, "libcuda lacks the VMM entry points");
    g_load_status = -2;
    return;
  }
  CUresult r = g_drv.Init(0);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof g_last_error, "cuInit failed (%d)", (int)r);
    g_load_status = -3;
    return;
  }
  g_load_status = 0;
}

int fail(CUresult r, const char* what) {
  const char* msg = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &msg);
  snprintf(g_last_error, sizeof g_last_error, "%s: %s (%d)", what, msg ? msg : "?", (int)r);
  return (int)r ? (int)r : -100;
}

#define DRV(call, what)                      \
  do {                                       \
    CUresult r_ = (call);                    \
    if (r_ != CUDA_SUCCESS) return fail(r_, what); \
  } while (0)

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int ensure_ctx(int device) {
  CUdevice dev;
  DRV(g_drv.DeviceGet(&dev, device), "cuDeviceGet");
  CUcontext ctx;
  DRV(g_drv.DevicePrimaryCtxRetain(&ctx, dev), "cuDevicePrimaryCtxRetain");
  DRV(g_drv.CtxSetCurrent(ctx), "cuCtxSetCurrent");
  return 0;
}

}  // namespace

extern "C" {

const char* adl_symm_last_error() { return g_last_error; }

// 0 on success. Safe to call repeatedly.
int adl_symm_init() {
  std::call_once(g_once, load_driver);
  return g_load_status;
}

// Topology probe for one device pair (peer access + native atomics + NVLink
// performance rank), and multicast support of a device.
int adl_topo_can_access_peer(int dev, int peer, int* can_access, int* atomics, int* perf_rank) {
  if (adl_symm_init() != 0) return -1;
  CUdevice a, b;
  DRV(g_drv.DeviceGet(&a, dev), "cuDeviceGet");
  DRV(g_drv.DeviceGet(&b, peer), "cuDeviceGet");
  DRV(g_drv.DeviceCanAccessPeer(can_access, a, b), "cuDeviceCanAccessPeer");
  *atomics = 0; *perf_rank = -1;
  if (*can_access) {
    g_drv.DeviceGetP2PAttribute(atomics, CU_DEVICE_P2P_ATTRIBUTE_NATIVE_ATOMIC_SUPPORTED, a, b);
    g_drv.DeviceGetP2PAttribute(perf_rank, CU_DEVICE_P2P_ATTRIBUTE_PERFORMANCE_RANK, a, b);
  }
  return 0;
}

int adl_topo_multicast_supported(int device) {
  if (adl_symm_init() != 0 || !g_drv.MulticastCreate) return 0;
  CUdevice dev;
  if (g_drv.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
  int v = 0;
  if (g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return 0;
  return v;
}

int adl_topo_vmm_fd_supported(int device) {
  if (adl_symm_init() != 0) return 0;
  CUdevice dev;
  if (g_drv.DeviceGet(&dev, device) != CUDA_SUCCESS) return 0;
  int v = 0;
  g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  return v;
}

// Round nbytes up to the allocation granularity of `device`.
int adl_symm_round_size(int device, size_t nbytes, size_t* out) {
  if (adl_symm_init() != 0) return -1;
  CUmemAllocationProp prop = alloc_prop(device);
  size_t gran = 0;
  DRV(g_drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
      "cuMemGetAllocationGranularity");
  *out = (nbytes + gran - 1) / gran * gran;
  return 0;
}

// Create a physical allocation on `device` and export it as a POSIX fd.
int adl_symm_create(int device, size_t size, unsigned long long* out_handle, int* out_fd) {
  if (adl_symm_init() != 0) return -1;
  if (int rc = ensure_ctx(device)) return rc;
  CUmemAllocationProp prop = alloc_prop(device);
  CUmemGenericAllocationHandle h;
  DRV(g_drv.MemCreate(&h, size, &prop, 0), "cuMemCreate");
  int fd = -1;
  CUresult r = g_drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    g_drv.MemRelease(h);
    return fail(r, "cuMemExportToShareableHandle");
  }
  *out_handle = (unsigned long long)h;
  *out_fd = fd;
  return 0;
}

// Import a peer's allocation from an fd received over a unix socket.
int adl_symm_import(int fd, unsigned long long* out_handle) {
  if (adl_symm_init() != 0) return -1;
  CUmemGenericAllocationHandle h;
  DRV(g_drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd,
                                         CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
      "cuMemImportFromShareableHandle");
  *out_handle = (unsigned long long)h;
  return 0;
}

// Map an allocation (own or imported) read/write for `device`.
int adl_symm_map(unsigned long long handle, size_t size, int device, unsigned long long* out_ptr) {
  if (adl_symm_init() != 0) return -1;
  if (int rc = ensure_ctx(device)) return rc;
  CUdeviceptr ptr = 0;
  DRV(g_drv.MemAddressReserve(&ptr, size, 0, 0, 0), "cuMemAddressReserve");
  CUresult r = g_drv.MemMap(ptr, size, 0, (CUmemGenericAllocationHandle)handle, 0);
  if (r != CUDA_SUCCESS) {
    g_drv.MemAddressFree(ptr, size);
    return fail(r, "cuMemMap");
  }
  CUmemAccessDesc desc;
  memset(&desc, 0, sizeof desc);
  desc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  desc.location.id = device;
  desc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = g_drv.MemSetAccess(ptr, size, &desc, 1);
  if (r != CUDA_SUCCESS) {
    g_drv.MemUnmap(ptr, size);
    g_drv.MemAddressFree(ptr, size);
    return fail(r, "cuMemSetAccess");
  }
  *out_ptr = (unsigned long long)ptr;
  return 0;
}

int adl_symm_unmap(unsigned long long ptr, size_t size) {
  if (adl_symm_init() != 0) return -1;
  DRV(g_drv.MemUnmap((CUdeviceptr)ptr, size), "cuMemUnmap");
  DRV(g_drv.MemAddressFree((CUdeviceptr)ptr, size), "cuMemAddressFree");
  return 0;
}

int adl_symm_release(unsigned long long handle) {
  if (adl_symm_init() != 0) return -1;
  DRV(g_drv.MemRelease((CUmemGenericAllocationHandle)handle), "cuMemRelease");
  return 0;
}

// ---- NVLS multicast --------------------------------------------------------

static CUmulticastObjectProp mc_prop(int num_devices, size_t size) {
  CUmulticastObjectProp prop;
  memset(&prop, 0, sizeof prop);
  prop.numDevices = (unsigned)num_devices;
  prop.size = size;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int adl_mc_round_size(int num_devices, size_t nbytes, size_t* out) {
  if (adl_symm_init() != 0 || !g_drv.MulticastGetGranularity) return -1;
  CUmulticastObjectProp prop = mc_prop(num_devices, nbytes);
  size_t gran = 0;
  DRV(g_drv.MulticastGetGranularity(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED),
      "cuMulticastGetGranularity");
  *out = (nbytes + gran - 1) / gran * gran;
  return 0;
}

// Rank 0: create the multicast object and export it.
int adl_mc_create(int num_devices, size_t size, unsigned long long* out_handle, int* out_fd) {
  if (adl_symm_init() != 0 || !g_drv.MulticastCreate) return -1;
  CUmulticastObjectProp prop = mc_prop(num_devices, size);
  CUmemGenericAllocationHandle h;
  DRV(g_drv.MulticastCreate(&h, &prop), "cuMulticastCreate");
  int fd = -1;
  CUresult r = g_drv.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    g_drv.MemRelease(h);
    return fail(r, "cuMemExportToShareableHandle(multicast)");
  }
  *out_handle = (unsigned long long)h;
  *out_fd = fd;
  return 0;
}

int adl_mc_add_device(unsigned long long mc_handle, int device) {
  if (adl_symm_init() != 0 || !g_drv.MulticastAddDevice) return -1;
  CUdevice dev;
  DRV(g_drv.DeviceGet(&dev, device), "cuDeviceGet");
  DRV(g_drv.MulticastAddDevice((CUmemGenericAllocationHandle)mc_handle, dev), "cuMulticastAddDevice");
  return 0;
}

// Bind this rank's physical allocation at offset 0 of the multicast object
// (call after EVERY rank has added its device).
int adl_mc_bind(unsigned long long mc_handle, unsigned long long mem_handle, size_t size) {
  if (adl_symm_init() != 0 || !g_drv.MulticastBindMem) return -1;
  DRV(g_drv.MulticastBindMem((CUmemGenericAllocationHandle)mc_handle, 0,
                             (CUmemGenericAllocationHandle)mem_handle, 0, size, 0),
      "cuMulticastBindMem");
  return 0;
}

}  // extern "C"
