#include "tfrecord.h"

#include <pybind11/stl.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
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
}

}  // namespace tfos
