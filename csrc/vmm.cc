// Symmetric memory on the CUDA virtual-memory-management API + NVLink-switch multicast (NVLS).
//
// the fused collectives address every rank's buffers from inside
// one kernel.  The CUDA-IPC path (binding.cpp: symm_alloc / symm_open) gives per-peer unicast
// pointers; this file adds the substrate for the in-switch path:
//
//   vmm_alloc      physical allocation (cuMemCreate, POSIX-fd shareable) mapped at a local VA
//   vmm_import     map a peer's allocation from its fd -> unicast peer pointer (P2P loads/stores)
//   mc_create      rank 0: cuMulticastCreate over `world` devices, exported as an fd
//   mc_import      other ranks: import the multicast object
//   mc_add_device  every rank adds its own device (must complete on all ranks before binding)
//   mc_bind        every rank binds its physical allocation at offset 0
//   mc_map         map the multicast object -> a VA on which `multimem.ld_reduce` reduces the
//                  same offset of every rank's buffer in the switch and `multimem.st` stores to
//                  all of them (csrc/optim_comm.cu, MULTIMEM variants)
//
// The file descriptors travel between the rank processes over an AF_UNIX socket with SCM_RIGHTS
// (tensorflowonspark_b200/parallel/fdshare.py); what to fetch from whom is published on the same
// control channel as the IPC handles (torch.distributed or the reservation board).
//
// The reference has no counterpart: it hands collectives to TensorFlow/NCCL
// (tensorflowonspark/TFSparkNode.py:373-384 only writes the cluster spec).
#include <cuda.h>
#include <cuda_runtime.h>
#include <pybind11/pybind11.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "vmm.h"

namespace py = pybind11;

namespace tfos {
namespace {

template <typename Fn>
Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr)
    throw std::runtime_error(std::string("CUDA driver entry point not available: ") + name);
  return reinterpret_cast<Fn>(p);
}

#define DRV(name) static auto p_##name = driver_fn<decltype(&name)>(#name)

void ck(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  static auto p_err = driver_fn<CUresult (*)(CUresult, const char**)>("cuGetErrorString");
  p_err(r, &s);
  throw std::runtime_error(std::string("tfos vmm: ") + what + " failed: " + (s ? s : "?") + " (" +
                           std::to_string(static_cast<int>(r)) + ")");
}

int cur_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) throw std::runtime_error("cudaGetDevice failed");
  cudaFree(nullptr);  // make sure the primary context exists
  return dev;
}

struct Mapping {
  CUmemGenericAllocationHandle handle;
  CUdeviceptr va;
  size_t size;
  bool multicast;
};
std::mutex g_mu;
std::map<uint64_t, Mapping> g_maps;  // keyed by VA
std::map<uint64_t, CUmemGenericAllocationHandle> g_mc;  // multicast objects by id
uint64_t g_next_mc = 1;

CUmemAllocationProp alloc_prop(int dev) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

size_t round_up(size_t n, size_t g) { return (n + g - 1) / g * g; }

CUdeviceptr map_handle(CUmemGenericAllocationHandle h, size_t size, size_t align, int dev) {
  DRV(cuMemAddressReserve);
  DRV(cuMemMap);
  DRV(cuMemSetAccess);
  CUdeviceptr va = 0;
  ck(p_cuMemAddressReserve(&va, size, align, 0, 0), "cuMemAddressReserve");
  ck(p_cuMemMap(va, size, 0, h, 0), "cuMemMap");
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ck(p_cuMemSetAccess(va, size, &acc, 1), "cuMemSetAccess");
  return va;
}

// {multicast: bool, granularity: bytes for allocations, mc_granularity: bytes for multicast}
py::dict vmm_info(int world) {
  DRV(cuDeviceGet);
  DRV(cuDeviceGetAttribute);
  DRV(cuMemGetAllocationGranularity);
  const int dev = cur_device();
  CUdevice cu;
  ck(p_cuDeviceGet(&cu, dev), "cuDeviceGet");
  int mc = 0, vmm = 0, fd_ok = 0;
  p_cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cu);
  p_cuDeviceGetAttribute(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cu);
  p_cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cu);
  py::dict out;
  out["vmm"] = vmm != 0 && fd_ok != 0;
  out["multicast"] = mc != 0;
  size_t g = 0;
  CUmemAllocationProp prop = alloc_prop(dev);
  if (vmm) ck(p_cuMemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
              "cuMemGetAllocationGranularity");
  out["granularity"] = g;
  size_t mg = 0;
  if (mc && world > 1) {
    DRV(cuMulticastGetGranularity);
    CUmulticastObjectProp mp = {};
    mp.numDevices = world;
    mp.size = g ? g : (2u << 20);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS)
      mg = 0;
  }
  out["mc_granularity"] = mg;
  return out;
}

// -> (va, size, fd): zero-filled physical allocation mapped read/write on the current device
py::tuple vmm_alloc(size_t bytes, size_t granularity) {
  DRV(cuMemCreate);
  DRV(cuMemExportToShareableHandle);
  const int dev = cur_device();
  CUmemAllocationProp prop = alloc_prop(dev);
  const size_t size = round_up(bytes < 1 ? 1 : bytes, granularity);
  CUmemGenericAllocationHandle h;
  ck(p_cuMemCreate(&h, size, &prop, 0), "cuMemCreate");
  CUdeviceptr va = map_handle(h, size, granularity, dev);
  if (cudaMemset(reinterpret_cast<void*>(va), 0, size) != cudaSuccess)
    throw std::runtime_error("tfos vmm: memset of a fresh allocation failed");
  cudaDeviceSynchronize();
  int fd = -1;
  ck(p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
     "cuMemExportToShareableHandle");
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[va] = Mapping{h, va, size, false};
  }
  return py::make_tuple(static_cast<uint64_t>(va), size, fd);
}

// map a peer's allocation (fd received over SCM_RIGHTS) -> unicast VA on the current device
uint64_t vmm_import(int fd, size_t size, size_t granularity) {
  DRV(cuMemImportFromShareableHandle);
  const int dev = cur_device();
  CUmemGenericAllocationHandle h;
  ck(p_cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                      CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
     "cuMemImportFromShareableHandle");
  ::close(fd);
  CUdeviceptr va = map_handle(h, size, granularity, dev);
  std::lock_guard<std::mutex> lk(g_mu);
  g_maps[va] = Mapping{h, va, size, false};
  return static_cast<uint64_t>(va);
}

void vmm_free(uint64_t va) {
  DRV(cuMemUnmap);
  DRV(cuMemRelease);
  DRV(cuMemAddressFree);
  Mapping m;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(va);
    if (it == g_maps.end()) return;
    m = it->second;
    g_maps.erase(it);
  }
  p_cuMemUnmap(m.va, m.size);
  p_cuMemAddressFree(m.va, m.size);
  if (!m.multicast) p_cuMemRelease(m.handle);
}

// rank 0 -> (mc_id, fd)
py::tuple mc_create(int world, size_t size) {
  DRV(cuMulticastCreate);
  DRV(cuMemExportToShareableHandle);
  cur_device();
  CUmulticastObjectProp mp = {};
  mp.numDevices = world;
  mp.size = size;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemGenericAllocationHandle mc;
  ck(p_cuMulticastCreate(&mc, &mp), "cuMulticastCreate");
  int fd = -1;
  ck(p_cuMemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
     "cuMemExportToShareableHandle(multicast)");
  std::lock_guard<std::mutex> lk(g_mu);
  const uint64_t id = g_next_mc++;
  g_mc[id] = mc;
  return py::make_tuple(id, fd);
}

uint64_t mc_import(int fd) {
  DRV(cuMemImportFromShareableHandle);
  cur_device();
  CUmemGenericAllocationHandle mc;
  ck(p_cuMemImportFromShareableHandle(&mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                      CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
     "cuMemImportFromShareableHandle(multicast)");
  ::close(fd);
  std::lock_guard<std::mutex> lk(g_mu);
  const uint64_t id = g_next_mc++;
  g_mc[id] = mc;
  return id;
}

CUmemGenericAllocationHandle mc_of(uint64_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_mc.find(id);
  if (it == g_mc.end()) throw std::runtime_error("tfos vmm: unknown multicast object");
  return it->second;
}

void mc_add_device(uint64_t id) {
  DRV(cuMulticastAddDevice);
  DRV(cuDeviceGet);
  CUdevice cu;
  ck(p_cuDeviceGet(&cu, cur_device()), "cuDeviceGet");
  ck(p_cuMulticastAddDevice(mc_of(id), cu), "cuMulticastAddDevice");
}

// bind the physical allocation mapped at `va` into the multicast object at offset 0
void mc_bind(uint64_t id, uint64_t va, size_t size) {
  DRV(cuMulticastBindMem);
  CUmemGenericAllocationHandle mem;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(va);
    if (it == g_maps.end()) throw std::runtime_error("tfos vmm: mc_bind of an unknown allocation");
    mem = it->second.handle;
  }
  ck(p_cuMulticastBindMem(mc_of(id), 0, mem, 0, size, 0), "cuMulticastBindMem");
}

uint64_t mc_map(uint64_t id, size_t size, size_t granularity) {
  CUmemGenericAllocationHandle mc = mc_of(id);
  CUdeviceptr va = map_handle(mc, size, granularity, cur_device());
  std::lock_guard<std::mutex> lk(g_mu);
  g_maps[va] = Mapping{mc, va, size, true};
  return static_cast<uint64_t>(va);
}

}  // namespace

void bind_vmm(py::module_& m) {
  m.def("vmm_info", &vmm_info, py::arg("world") = 1);
  m.def("vmm_alloc", &vmm_alloc);
  m.def("vmm_import", &vmm_import);
  m.def("vmm_free", &vmm_free);
  m.def("mc_create", &mc_create);
  m.def("mc_import", &mc_import);
  m.def("mc_add_device", &mc_add_device);
  m.def("mc_bind", &mc_bind);
  m.def("mc_map", &mc_map);
}

}  // namespace tfos
