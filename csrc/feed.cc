#include "feed.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <pybind11/stl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>

namespace py = pybind11;

namespace tfos {
namespace {

constexpr uint32_t kMagic = 0x74664f53;  // "tfOS"

struct SlotMeta {
  std::atomic<uint64_t> seq;
  uint64_t nbytes;
  int64_t tag;
  uint32_t nrows;
  uint32_t pad;
};

struct Header {
  uint32_t magic;
  uint32_t nslots;
  uint64_t slot_bytes;
  uint64_t data_offset;
  std::atomic<uint64_t> head;  // next write position
  std::atomic<uint64_t> tail;  // next read position
  std::atomic<uint32_t> closed;
  uint32_t pad;
};

static_assert(std::atomic<uint64_t>::is_always_lock_free, "need address-free 64-bit atomics");

// Bounded multi-producer / multi-consumer queue (per-slot sequence numbers)
// whose payload slots live in the same shared mapping.
class ShmRing {
 public:
  ShmRing(const std::string& name, bool create, uint32_t nslots, uint64_t slot_bytes)
      : name_(name), owner_(create) {
    const uint64_t meta = sizeof(Header) + sizeof(SlotMeta) * (create ? nslots : 0);
    int fd;
    if (create) {
      shm_unlink(name.c_str());
      fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0) throw std::runtime_error("shm_open(create) failed for " + name);
      const uint64_t data_off = (meta + 4095) & ~4095ull;
      slot_bytes = (slot_bytes + 4095) & ~4095ull;
      size_ = data_off + slot_bytes * nslots;
      if (ftruncate(fd, static_cast<off_t>(size_)) != 0) {
        close(fd);
        throw std::runtime_error("ftruncate failed for " + name);
      }
      base_ = static_cast<uint8_t*>(mmap(nullptr, size_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
      close(fd);
      if (base_ == MAP_FAILED) throw std::runtime_error("mmap failed for " + name);
      hdr_ = reinterpret_cast<Header*>(base_);
      hdr_->nslots = nslots;
      hdr_->slot_bytes = slot_bytes;
      hdr_->data_offset = data_off;
      hdr_->head.store(0);
      hdr_->tail.store(0);
      hdr_->closed.store(0);
      slots_ = reinterpret_cast<SlotMeta*>(base_ + sizeof(Header));
      for (uint32_t i = 0; i < nslots; ++i) slots_[i].seq.store(i);
      std::atomic_thread_fence(std::memory_order_release);
      hdr_->magic = kMagic;
      // First-touch page faults of a fresh tmpfs mapping cost ~0.5 s per 64 MiB here - far more
      // than the memcpy that fills a slot.  Populate the payload pages in the background
      // (MADV_POPULATE_WRITE allocates without modifying contents, so it may race with feeders).
      // Detached, with a shared stop flag: the owner may be destroyed - or the process forked
      // (node processes are forked from the executor that created the ring; a std::thread copied
      // into a child must never be joined) - while the population is still running.
      uint8_t* p = base_ + data_off;
      const uint64_t n = slot_bytes * nslots;
      stop_populate_ = std::make_shared<std::atomic<bool>>(false);
      std::shared_ptr<std::atomic<bool>> stop = stop_populate_;
      std::thread([p, n, stop]() {
#ifdef MADV_POPULATE_WRITE
        const uint64_t step = 4ull << 20;
        for (uint64_t o = 0; o < n && !stop->load(); o += step)
          if (madvise(p + o, (n - o < step) ? n - o : step, MADV_POPULATE_WRITE) != 0) break;
#else
        (void)p;
        (void)n;
#endif
      }).detach();
    } else {
      fd = shm_open(name.c_str(), O_RDWR, 0600);
      if (fd < 0) throw std::runtime_error("shm_open(attach) failed for " + name);
      struct stat st;
      if (fstat(fd, &st) != 0) {
        close(fd);
        throw std::runtime_error("fstat failed for " + name);
      }
      size_ = static_cast<uint64_t>(st.st_size);
      base_ = static_cast<uint8_t*>(mmap(nullptr, size_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
      close(fd);
      if (base_ == MAP_FAILED) throw std::runtime_error("mmap failed for " + name);
      hdr_ = reinterpret_cast<Header*>(base_);
      if (hdr_->magic != kMagic) throw std::runtime_error("not a tfos ring: " + name);
      slots_ = reinterpret_cast<SlotMeta*>(base_ + sizeof(Header));
    }
  }
  ~ShmRing() {
    if (stop_populate_) stop_populate_->store(true);
    if (pinned_) cudaHostUnregister(base_);
    if (base_ && base_ != MAP_FAILED) munmap(base_, size_);
    if (owner_) shm_unlink(name_.c_str());
  }

  uint32_t nslots() const { return hdr_->nslots; }
  uint64_t slot_bytes() const { return hdr_->slot_bytes; }
  uint8_t* slot_ptr(uint32_t i) const { return base_ + hdr_->data_offset + hdr_->slot_bytes * i; }
  void close_ring() { hdr_->closed.store(1); }
  bool closed() const { return hdr_->closed.load() != 0; }
  uint64_t depth() const { return hdr_->head.load() - hdr_->tail.load(); }

  // returns position, or -1 on timeout / closed
  int64_t acquire_write(double timeout_s) {
    const auto deadline = now() + timeout_s;
    for (;;) {
      uint64_t pos = hdr_->head.load(std::memory_order_relaxed);
      SlotMeta& s = slots_[pos % hdr_->nslots];
      const uint64_t seq = s.seq.load(std::memory_order_acquire);
      const int64_t dif = static_cast<int64_t>(seq) - static_cast<int64_t>(pos);
      if (dif == 0) {
        if (hdr_->head.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed))
          return static_cast<int64_t>(pos);
      } else if (dif < 0) {  // full
        if (closed() || now() > deadline) return -1;
        nap();
      }
    }
  }
  void commit_write(int64_t pos, uint64_t nbytes, uint32_t nrows, int64_t tag) {
    SlotMeta& s = slots_[static_cast<uint64_t>(pos) % hdr_->nslots];
    s.nbytes = nbytes;
    s.nrows = nrows;
    s.tag = tag;
    s.seq.store(static_cast<uint64_t>(pos) + 1, std::memory_order_release);
  }
  int64_t acquire_read(double timeout_s) {
    const auto deadline = now() + timeout_s;
    for (;;) {
      uint64_t pos = hdr_->tail.load(std::memory_order_relaxed);
      SlotMeta& s = slots_[pos % hdr_->nslots];
      const uint64_t seq = s.seq.load(std::memory_order_acquire);
      const int64_t dif = static_cast<int64_t>(seq) - static_cast<int64_t>(pos + 1);
      if (dif == 0) {
        if (hdr_->tail.compare_exchange_weak(pos, pos + 1, std::memory_order_relaxed))
          return static_cast<int64_t>(pos);
      } else if (dif < 0) {  // empty
        if (now() > deadline) return -1;
        if (closed() && hdr_->head.load() == pos) return -2;
        nap();
      }
    }
  }
  void release_read(int64_t pos) {
    SlotMeta& s = slots_[static_cast<uint64_t>(pos) % hdr_->nslots];
    s.seq.store(static_cast<uint64_t>(pos) + hdr_->nslots, std::memory_order_release);
  }
  const SlotMeta& meta(int64_t pos) const { return slots_[static_cast<uint64_t>(pos) % hdr_->nslots]; }

  // Page-lock the whole mapping so cudaMemcpyAsync from a slot is a true DMA.
  void pin() {
    if (pinned_) return;
    cudaError_t e = cudaHostRegister(base_, size_, cudaHostRegisterPortable);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("cudaHostRegister failed: ") + cudaGetErrorString(e));
    pinned_ = true;
  }
  bool pinned() const { return pinned_; }
  void h2d(int64_t pos, uint64_t src_off, uint64_t dst_ptr, uint64_t nbytes, uint64_t stream) {
    const uint8_t* src = slot_ptr(static_cast<uint64_t>(pos) % hdr_->nslots) + src_off;
    cudaError_t e = cudaMemcpyAsync(reinterpret_cast<void*>(dst_ptr), src, nbytes,
                                    cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("cudaMemcpyAsync failed: ") + cudaGetErrorString(e));
  }

 private:
  static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  static void nap() { std::this_thread::sleep_for(std::chrono::microseconds(50)); }

  std::string name_;
  bool owner_;
  uint8_t* base_ = nullptr;
  uint64_t size_ = 0;
  Header* hdr_ = nullptr;
  SlotMeta* slots_ = nullptr;
  bool pinned_ = false;
  std::shared_ptr<std::atomic<bool>> stop_populate_;
};

}  // namespace

void bind_feed(py::module_& m) {
  py::class_<ShmRing>(m, "ShmRing")
      .def(py::init<const std::string&, bool, uint32_t, uint64_t>(), py::arg("name"),
           py::arg("create"), py::arg("nslots") = 0, py::arg("slot_bytes") = 0)
      .def_property_readonly("nslots", &ShmRing::nslots)
      .def_property_readonly("slot_bytes", &ShmRing::slot_bytes)
      .def_property_readonly("pinned", &ShmRing::pinned)
      .def("depth", &ShmRing::depth)
      .def("close", &ShmRing::close_ring)
      .def("closed", &ShmRing::closed)
      .def("acquire_write",
           [](ShmRing& r, double t) {
             py::gil_scoped_release rel;
             return r.acquire_write(t);
           })
      .def("commit_write", &ShmRing::commit_write, py::arg("pos"), py::arg("nbytes"),
           py::arg("nrows"), py::arg("tag") = 0)
      .def("acquire_read",
           [](ShmRing& r, double t) {
             py::gil_scoped_release rel;
             return r.acquire_read(t);
           })
      .def("release_read", &ShmRing::release_read)
      .def("meta",
           [](ShmRing& r, int64_t pos) {
             const SlotMeta& s = r.meta(pos);
             return py::make_tuple(s.nbytes, s.nrows, s.tag);
           })
      .def("slot_view",
           [](ShmRing& r, int64_t pos) {
             return py::memoryview::from_memory(
                 r.slot_ptr(static_cast<uint64_t>(pos) % r.nslots()),
                 static_cast<py::ssize_t>(r.slot_bytes()), /*readonly=*/false);
           },
           py::keep_alive<0, 1>())
      .def("pin", &ShmRing::pin)
      .def("h2d", &ShmRing::h2d, py::arg("pos"), py::arg("src_off"), py::arg("dst_ptr"),
           py::arg("nbytes"), py::arg("stream") = 0);
}

}  // namespace tfos
