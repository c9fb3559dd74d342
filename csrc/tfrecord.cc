#include "tfrecord.h"

#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace py = pybind11;

namespace tfos {
namespace {

// ------------------------------------------------------------------ CRC32C
uint32_t g_table[8][256];
bool g_table_ready = false;

void init_table() {
  if (g_table_ready) return;
  const uint32_t poly = 0x82f63b78u;  // reflected Castagnoli
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ poly : c >> 1;
    g_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_table[t][i] = (g_table[t - 1][i] >> 8) ^ g_table[0][g_table[t - 1][i] & 0xff];
  g_table_ready = true;
}

uint32_t crc32c(const uint8_t* p, size_t n) {
  init_table();
  uint32_t c = 0xffffffffu;
  while (n >= 8) {  // slicing-by-8
    uint64_t v;
    std::memcpy(&v, p, 8);
    v ^= c;
    c = g_table[7][v & 0xff] ^ g_table[6][(v >> 8) & 0xff] ^ g_table[5][(v >> 16) & 0xff] ^
        g_table[4][(v >> 24) & 0xff] ^ g_table[3][(v >> 32) & 0xff] ^ g_table[2][(v >> 40) & 0xff] ^
        g_table[1][(v >> 48) & 0xff] ^ g_table[0][(v >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

uint32_t masked(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

// --------------------------------------------------------- record framing
// uint64 length | uint32 masked_crc(length) | bytes | uint32 masked_crc(bytes)
void write_records(const std::string& path, const std::vector<py::bytes>& records, bool append) {
  FILE* f = std::fopen(path.c_str(), append ? "ab" : "wb");
  if (!f) throw std::runtime_error("cannot open " + path + " for writing");
  for (const auto& r : records) {
    std::string s = r;
    uint64_t len = s.size();
    uint8_t hdr[12];
    std::memcpy(hdr, &len, 8);
    uint32_t c = masked(crc32c(hdr, 8));
    std::memcpy(hdr + 8, &c, 4);
    uint32_t d = masked(crc32c(reinterpret_cast<const uint8_t*>(s.data()), s.size()));
    if (std::fwrite(hdr, 1, 12, f) != 12 || std::fwrite(s.data(), 1, s.size(), f) != s.size() ||
        std::fwrite(&d, 1, 4, f) != 4) {
      std::fclose(f);
      throw std::runtime_error("short write to " + path);
    }
  }
  std::fclose(f);
}

std::vector<py::bytes> read_records(const std::string& path, bool verify) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<py::bytes> out;
  std::string buf;
  for (;;) {
    uint8_t hdr[12];
    size_t got = std::fread(hdr, 1, 12, f);
    if (got == 0) break;
    if (got != 12) {
      std::fclose(f);
      throw std::runtime_error("truncated TFRecord header in " + path);
    }
    uint64_t len;
    uint32_t c;
    std::memcpy(&len, hdr, 8);
    std::memcpy(&c, hdr + 8, 4);
    if (verify && masked(crc32c(hdr, 8)) != c) {
      std::fclose(f);
      throw std::runtime_error("corrupt TFRecord length CRC in " + path);
    }
    buf.resize(len);
    uint32_t d;
    if (std::fread(&buf[0], 1, len, f) != len || std::fread(&d, 1, 4, f) != 4) {
      std::fclose(f);
      throw std::runtime_error("truncated TFRecord payload in " + path);
    }
    if (verify && masked(crc32c(reinterpret_cast<const uint8_t*>(buf.data()), len)) != d) {
      std::fclose(f);
      throw std::runtime_error("corrupt TFRecord data CRC in " + path);
    }
    out.emplace_back(buf);
  }
  std::fclose(f);
  return out;
}

// ------------------------------------------------------ protobuf wire codec
void put_varint(std::string& o, uint64_t v) {
  while (v >= 0x80) {
    o.push_back(static_cast<char>((v & 0x7f) | 0x80));
    v >>= 7;
  }
  o.push_back(static_cast<char>(v));
}
void put_len(std::string& o, int field, const std::string& payload) {
  put_varint(o, (static_cast<uint64_t>(field) << 3) | 2);
  put_varint(o, payload.size());
  o += payload;
}

struct Reader {
  const uint8_t* p;
  const uint8_t* e;
  bool done() const { return p >= e; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < e) {
      uint8_t b = *p++;
      v |= static_cast<uint64_t>(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
      if (shift > 63) break;
    }
    throw std::runtime_error("bad varint in Example");
  }
  Reader sub() {
    uint64_t n = varint();
    if (n > static_cast<uint64_t>(e - p)) throw std::runtime_error("bad length in Example");
    Reader r{p, p + n};
    p += n;
    return r;
  }
  // advance over a fixed-width field; a record that ends inside it is malformed
  const uint8_t* fixed(int n) {
    if (static_cast<int64_t>(e - p) < n) throw std::runtime_error("truncated fixed-width field in Example");
    const uint8_t* q = p;
    p += n;
    return q;
  }
  void skip(int wt) {
    switch (wt) {
      case 0: varint(); break;
      case 1: fixed(8); break;
      case 2: sub(); break;
      case 5: fixed(4); break;
      default: throw std::runtime_error("unsupported wire type in Example");
    }
  }
};

// features: dict name -> (kind, list) with kind in {"bytes", "float", "int64"}
py::bytes encode_example(const py::dict& features) {
  std::string feats;
  for (auto item : features) {
    const std::string name = py::cast<std::string>(item.first);
    py::tuple kv = py::cast<py::tuple>(item.second);
    const std::string kind = py::cast<std::string>(kv[0]);
    py::list vals = py::cast<py::list>(kv[1]);
    std::string lst, feature;
    if (kind == "bytes") {
      for (auto v : vals) put_len(lst, 1, py::cast<std::string>(py::cast<py::bytes>(v)));
      put_len(feature, 1, lst);
    } else if (kind == "float") {
      std::string packed;
      for (auto v : vals) {
        float f = py::cast<float>(v);
        packed.append(reinterpret_cast<const char*>(&f), 4);
      }
      if (!packed.empty()) put_len(lst, 1, packed);
      put_len(feature, 2, lst);
    } else if (kind == "int64") {
      std::string packed;
      for (auto v : vals) put_varint(packed, static_cast<uint64_t>(py::cast<int64_t>(v)));
      if (!packed.empty()) put_len(lst, 1, packed);
      put_len(feature, 3, lst);
    } else {
      throw std::runtime_error("unknown feature kind " + kind);
    }
    std::string entry;
    put_len(entry, 1, name);
    put_len(entry, 2, feature);
    put_len(feats, 1, entry);
  }
  std::string example;
  put_len(example, 1, feats);
  return py::bytes(example);
}

py::dict decode_example(const py::bytes& data) {
  std::string s = data;
  Reader ex{reinterpret_cast<const uint8_t*>(s.data()),
            reinterpret_cast<const uint8_t*>(s.data()) + s.size()};
  py::dict out;
  while (!ex.done()) {
    uint64_t tag = ex.varint();
    if ((tag >> 3) != 1 || (tag & 7) != 2) {
      ex.skip(tag & 7);
      continue;
    }
    Reader feats = ex.sub();
    while (!feats.done()) {
      uint64_t t2 = feats.varint();
      if ((t2 >> 3) != 1 || (t2 & 7) != 2) {
        feats.skip(t2 & 7);
        continue;
      }
      Reader entry = feats.sub();
      std::string name;
      std::string kind = "bytes";
      py::list values;
      while (!entry.done()) {
        uint64_t t3 = entry.varint();
        const int field = static_cast<int>(t3 >> 3), wt = static_cast<int>(t3 & 7);
        if (field == 1 && wt == 2) {
          Reader k = entry.sub();
          name.assign(reinterpret_cast<const char*>(k.p), k.e - k.p);
        } else if (field == 2 && wt == 2) {
          Reader feature = entry.sub();
          while (!feature.done()) {
            uint64_t t4 = feature.varint();
            const int kf = static_cast<int>(t4 >> 3);
            if ((t4 & 7) != 2) {
              feature.skip(t4 & 7);
              continue;
            }
            Reader lst = feature.sub();
            kind = kf == 1 ? "bytes" : (kf == 2 ? "float" : "int64");
            while (!lst.done()) {
              uint64_t t5 = lst.varint();
              const int wt5 = static_cast<int>(t5 & 7);
              if ((t5 >> 3) != 1) {
                lst.skip(wt5);
                continue;
              }
              if (kf == 1) {
                if (wt5 != 2) throw std::runtime_error("BytesList value with a non-bytes wire type");
                Reader b = lst.sub();
                values.append(py::bytes(reinterpret_cast<const char*>(b.p), b.e - b.p));
              } else if (kf == 2) {
                if (wt5 == 2) {
                  Reader pk = lst.sub();
                  for (const uint8_t* q = pk.p; q + 4 <= pk.e; q += 4) {
                    float f;
                    std::memcpy(&f, q, 4);
                    values.append(f);
                  }
                } else if (wt5 == 5) {
                  float f;
                  std::memcpy(&f, lst.fixed(4), 4);
                  values.append(f);
                } else {
                  throw std::runtime_error("FloatList value with a non-float wire type");
                }
              } else {
                if (wt5 == 2) {
                  Reader pk = lst.sub();
                  while (!pk.done()) values.append(static_cast<int64_t>(pk.varint()));
                } else if (wt5 == 0) {
                  values.append(static_cast<int64_t>(lst.varint()));
                } else {
                  throw std::runtime_error("Int64List value with a non-varint wire type");
                }
              }
            }
          }
        } else {
          entry.skip(wt);
        }
      }
      out[py::str(name)] = py::make_tuple(kind, values);
    }
  }
  return out;
}

// ------------------------------------------------------ batch decode into dense arrays
// A training input pipeline wants [N, L] arrays, not N dicts of python lists: this walks N
// serialized Examples once, writes every requested feature straight into its row of a numpy
// array (int64 / int32 / uint8 for Int64List, float32 for FloatList, raw uint8 for a
// fixed-length BytesList value), with the GIL released and the records split over threads.
struct FeatSpec {
  std::string name;
  int kind;      // 1 bytes, 2 float, 3 int64 (the Feature oneof field numbers)
  int64_t len;   // values per record
  int out;       // element type of the output: 0 int64, 1 int32, 2 uint8, 3 float32
  int64_t esize;
  uint8_t* base;
};

inline void store_int(const FeatSpec& f, uint8_t* row, int64_t k, int64_t v) {
  switch (f.out) {
    case 0: reinterpret_cast<int64_t*>(row)[k] = v; break;
    case 1: reinterpret_cast<int32_t*>(row)[k] = static_cast<int32_t>(v); break;
    default: row[k] = static_cast<uint8_t>(v); break;
  }
}

// the hot loop of an image-like feature: a packed run of varints into a typed row.  Returns the
// new fill count, -1 when the run holds more than ``cap`` values, -2 on a malformed varint.
template <typename T>
inline int64_t unpack_varints(const uint8_t* p, const uint8_t* e, T* dst, int64_t k, int64_t cap) {
  // values below 2^14 (pixels, labels, small ids) are one or two bytes: decoded without a
  // data-dependent branch - for random pixel data the "is there a second byte" branch would
  // mispredict every other value
  while (e - p >= 2 && k < cap) {
    const uint32_t b0 = p[0], b1 = p[1];
    if (__builtin_expect((b0 & b1 & 0x80) != 0, 0)) break;   // three bytes or more: generic loop
    const uint32_t two = b0 >> 7;
    dst[k++] = static_cast<T>((b0 & 0x7f) | (((b1 & 0x7f) << 7) & (0u - two)));
    p += 1 + two;
  }
  while (p < e) {
    if (k >= cap) return -1;
    uint64_t v = *p++;
    if (v & 0x80) {
      v &= 0x7f;
      int shift = 7;
      for (;;) {
        if (p >= e || shift > 63) return -2;
        const uint8_t b = *p++;
        v |= static_cast<uint64_t>(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
      }
    }
    dst[k++] = static_cast<T>(v);
  }
  return k;
}

void decode_one(const uint8_t* p, size_t n, int64_t r, std::vector<FeatSpec>& spec) {
  Reader ex{p, p + n};
  std::vector<int64_t> seen(spec.size(), -1);
  while (!ex.done()) {
    uint64_t tag = ex.varint();
    if ((tag >> 3) != 1 || (tag & 7) != 2) {
      ex.skip(tag & 7);
      continue;
    }
    Reader feats = ex.sub();
    while (!feats.done()) {
      uint64_t t2 = feats.varint();
      if ((t2 >> 3) != 1 || (t2 & 7) != 2) {
        feats.skip(t2 & 7);
        continue;
      }
      Reader entry = feats.sub();
      int which = -1;
      Reader feature{nullptr, nullptr};
      while (!entry.done()) {
        uint64_t t3 = entry.varint();
        const int field = static_cast<int>(t3 >> 3), wt = static_cast<int>(t3 & 7);
        if (field == 1 && wt == 2) {
          Reader k = entry.sub();
          const size_t kl = k.e - k.p;
          for (size_t i = 0; i < spec.size(); ++i)
            if (spec[i].name.size() == kl && std::memcmp(spec[i].name.data(), k.p, kl) == 0) which = static_cast<int>(i);
        } else if (field == 2 && wt == 2) {
          feature = entry.sub();
        } else {
          entry.skip(wt);
        }
      }
      if (which < 0 || feature.p == nullptr) continue;   // a feature nobody asked for
      FeatSpec& f = spec[which];
      uint8_t* row = f.base + r * f.len * f.esize;
      int64_t k = 0;
      while (!feature.done()) {
        uint64_t t4 = feature.varint();
        const int kf = static_cast<int>(t4 >> 3);
        if ((t4 & 7) != 2) {
          feature.skip(t4 & 7);
          continue;
        }
        Reader lst = feature.sub();
        if (kf != f.kind) throw std::runtime_error("feature '" + f.name + "' has another type than requested");
        while (!lst.done()) {
          uint64_t t5 = lst.varint();
          const int wt5 = static_cast<int>(t5 & 7);
          if ((t5 >> 3) != 1) {
            lst.skip(wt5);
            continue;
          }
          if (kf == 1) {
            if (wt5 != 2) throw std::runtime_error("BytesList value with a non-bytes wire type");
            Reader b = lst.sub();
            if (k != 0 || (b.e - b.p) != f.len)
              throw std::runtime_error("bytes feature '" + f.name + "' is not one value of the requested length");
            std::memcpy(row, b.p, f.len);
            k = f.len;
          } else if (kf == 2) {
            if (wt5 == 2) {
              Reader pk = lst.sub();
              const int64_t cnt = (pk.e - pk.p) / 4;
              if (k + cnt > f.len) throw std::runtime_error("feature '" + f.name + "' holds more values than requested");
              std::memcpy(row + k * 4, pk.p, cnt * 4);
              k += cnt;
            } else if (wt5 == 5) {
              if (k + 1 > f.len) throw std::runtime_error("feature '" + f.name + "' holds more values than requested");
              std::memcpy(row + k * 4, lst.fixed(4), 4);
              ++k;
            } else {
              throw std::runtime_error("FloatList value with a non-float wire type");
            }
          } else {
            if (wt5 == 2) {
              Reader pk = lst.sub();
              int64_t nk;
              if (f.out == 0)
                nk = unpack_varints(pk.p, pk.e, reinterpret_cast<int64_t*>(row), k, f.len);
              else if (f.out == 1)
                nk = unpack_varints(pk.p, pk.e, reinterpret_cast<int32_t*>(row), k, f.len);
              else
                nk = unpack_varints(pk.p, pk.e, row, k, f.len);
              if (nk == -1) throw std::runtime_error("feature '" + f.name + "' holds more values than requested");
              if (nk < 0) throw std::runtime_error("bad varint in Example");
              k = nk;
            } else if (wt5 == 0) {
              if (k >= f.len) throw std::runtime_error("feature '" + f.name + "' holds more values than requested");
              store_int(f, row, k++, static_cast<int64_t>(lst.varint()));
            } else {
              throw std::runtime_error("Int64List value with a non-varint wire type");
            }
          }
        }
      }
      seen[which] = k;
    }
  }
  for (size_t i = 0; i < spec.size(); ++i)
    if (seen[i] != spec[i].len)
      throw std::runtime_error("record " + std::to_string(r) + ": feature '" + spec[i].name + "' has " +
                               (seen[i] < 0 ? std::string("no") : std::to_string(seen[i])) + " values, " +
                               std::to_string(spec[i].len) + " requested");
}

// spec: list of (name, kind, length, out_dtype) with kind in {"bytes","float","int64"} and
// out_dtype in {"int64","int32","uint8","float32"}; returns {name: ndarray [N, length]}
py::dict decode_batch(const std::vector<py::bytes>& records, const py::list& spec_in, int threads) {
  const int64_t n = static_cast<int64_t>(records.size());
  std::vector<FeatSpec> spec;
  py::dict out;
  for (auto item : spec_in) {
    py::tuple t = py::cast<py::tuple>(item);
    FeatSpec f;
    f.name = py::cast<std::string>(t[0]);
    const std::string kind = py::cast<std::string>(t[1]);
    f.len = py::cast<int64_t>(t[2]);
    const std::string od = py::cast<std::string>(t[3]);
    f.kind = kind == "bytes" ? 1 : (kind == "float" ? 2 : (kind == "int64" ? 3 : 0));
    if (f.kind == 0 || f.len < 0) throw std::runtime_error("bad feature spec for " + f.name);
    py::array arr;
    if (f.kind == 2) {
      if (od != "float32") throw std::runtime_error("float features decode to float32");
      f.out = 3, f.esize = 4;
      arr = py::array_t<float>({n, f.len});
    } else if (f.kind == 1) {
      if (od != "uint8") throw std::runtime_error("bytes features decode to uint8");
      f.out = 2, f.esize = 1;
      arr = py::array_t<uint8_t>({n, f.len});
    } else if (od == "int64") {
      f.out = 0, f.esize = 8;
      arr = py::array_t<int64_t>({n, f.len});
    } else if (od == "int32") {
      f.out = 1, f.esize = 4;
      arr = py::array_t<int32_t>({n, f.len});
    } else if (od == "uint8") {
      f.out = 2, f.esize = 1;
      arr = py::array_t<uint8_t>({n, f.len});
    } else {
      throw std::runtime_error("int64 features decode to int64, int32 or uint8");
    }
    f.base = static_cast<uint8_t*>(arr.mutable_data());
    out[py::str(f.name)] = arr;
    spec.push_back(f);
  }
  // borrow the record buffers while the GIL is held; the list keeps them alive
  std::vector<std::pair<const uint8_t*, size_t>> views(n);
  for (int64_t i = 0; i < n; ++i) {
    char* buf;
    Py_ssize_t len;
    if (PyBytes_AsStringAndSize(records[i].ptr(), &buf, &len) != 0) throw py::error_already_set();
    views[i] = {reinterpret_cast<const uint8_t*>(buf), static_cast<size_t>(len)};
  }
  std::string error;
  {
    py::gil_scoped_release nogil;
    int workers = threads > 0 ? threads : 1;
    if (workers > n / 64) workers = n >= 128 ? static_cast<int>(n / 64) : 1;   // >= 64 records per thread
    std::atomic<bool> failed{false};
    std::vector<std::string> errors(workers);
    auto work = [&](int w) {
      std::vector<FeatSpec> mine = spec;
      for (int64_t i = w; i < n && !failed.load(); i += workers) {
        try {
          decode_one(views[i].first, views[i].second, i, mine);
        } catch (const std::exception& e) {
          errors[w] = e.what();
          failed.store(true);
          return;
        }
      }
    };
    if (workers == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int w = 0; w < workers; ++w) pool.emplace_back(work, w);
      for (auto& t : pool) t.join();
    }
    for (const auto& e : errors)
      if (!e.empty() && error.empty()) error = e;
  }
  if (!error.empty()) throw std::runtime_error(error);
  return out;
}

}  // namespace

void bind_tfrecord(py::module_& m) {
  m.def("crc32c", [](const py::bytes& b) {
    std::string s = b;
    return crc32c(reinterpret_cast<const uint8_t*>(s.data()), s.size());
  });
  m.def("masked_crc32c", [](const py::bytes& b) {
    std::string s = b;
    return masked(crc32c(reinterpret_cast<const uint8_t*>(s.data()), s.size()));
  });
  m.def("tfrecord_write", &write_records, py::arg("path"), py::arg("records"),
        py::arg("append") = false);
  m.def("tfrecord_read", &read_records, py::arg("path"), py::arg("verify") = true);
  m.def("example_encode", &encode_example);
  m.def("example_decode", &decode_example);
  m.def("example_decode_batch", &decode_batch, py::arg("records"), py::arg("spec"), py::arg("threads") = 1);
}

}  // namespace tfos
