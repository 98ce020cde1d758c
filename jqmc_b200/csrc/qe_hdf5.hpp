// Minimal HDF5 reader for the engine's native input path: jQMC's `hamiltonian_data.h5` / `restart.h5`
// (groups, contiguous or compact numeric / string datasets, scalar attributes; jqmc/hamiltonians.py:369-573,
// jqmc/_checkpoint.py:1-30) without libhdf5, which this image does not have.
//
// Supported on-disk structures (HDF5 File Format Specification): superblock v0/v1 with 8-byte offsets and lengths, version-1
// object headers (with continuation blocks), version-1 group B-trees + local heaps + symbol-table nodes, dataspace v1/v2,
// fixed-point / IEEE float / fixed-length string datatypes, variable-length strings through the global heap (h5py writes
// str attributes that way), data layout v3 contiguous / compact and v1/v2 contiguous, attribute messages v1-v3.
// Anything else throws std::runtime_error: never a silent mis-read.  Same subset as jqmc_b200/hdf5_lite.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace qeio {

struct H5Value {
  // numeric data converted to double / int64 on request; strings kept as text
  std::vector<uint64_t> shape;
  int cls = -1;  // 0 fixed point, 1 float, 3 string (fixed or variable length)
  int size = 0;  // element size in bytes (fixed-point / float / fixed-length string)
  bool is_signed = true;
  std::vector<unsigned char> raw;       // numeric / fixed-length string payload
  std::vector<std::string> strings;     // decoded strings (cls 3)
  size_t count() const {
    size_t n = 1;
    for (uint64_t s : shape) n *= (size_t)s;
    return n;
  }
  std::vector<double> as_double() const {
    std::vector<double> out(count());
    for (size_t i = 0; i < out.size(); ++i) out[i] = get_double(i);
    return out;
  }
  std::vector<int32_t> as_int() const {
    std::vector<int32_t> out(count());
    for (size_t i = 0; i < out.size(); ++i) out[i] = (int32_t)get_double(i);
    return out;
  }
  double get_double(size_t i) const {
    const unsigned char* p = raw.data() + i * size;
    if (cls == 1) {
      if (size == 8) { double v; std::memcpy(&v, p, 8); return v; }
      if (size == 4) { float v; std::memcpy(&v, p, 4); return v; }
      throw std::runtime_error("hdf5: unsupported float size");
    }
    if (cls == 0) {
      if (size > 8) throw std::runtime_error("hdf5: integer wider than 64 bits");
      uint64_t u = 0;
      std::memcpy(&u, p, size);
      if (is_signed && size < 8 && (u >> (8 * size - 1))) u |= ~uint64_t(0) << (8 * size);
      return is_signed ? (double)(int64_t)u : (double)u;
    }
    throw std::runtime_error("hdf5: not a numeric value");
  }
};

class H5File {
 public:
  explicit H5File(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("hdf5: cannot open " + path);
    buf_.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf_.size() < 96 || std::memcmp(buf_.data(), sig, 8) != 0) throw std::runtime_error("hdf5: not an HDF5 file: " + path);
    const int ver = u8(8);
    if (ver != 0 && ver != 1) throw std::runtime_error("hdf5: superblock version " + std::to_string(ver) + " is not supported");
    if (u8(13) != 8 || u8(14) != 8) throw std::runtime_error("hdf5: only 8-byte offsets / lengths");
    size_t p = ver == 0 ? 24 : 28;
    base_ = u64(p);
    p += 32;
    root_ = u64(p + 8);  // object header address of the root symbol-table entry
  }

  bool has(const std::string& path) const {
    try { resolve(path); return true; } catch (const std::out_of_range&) { return false; }
  }
  bool is_group(const std::string& path) const {
    std::map<std::string, uint64_t> ch;
    return children(resolve(path), ch);
  }
  std::vector<std::string> list(const std::string& path) const {
    std::map<std::string, uint64_t> ch;
    if (!children(resolve(path), ch)) throw std::runtime_error("hdf5: " + path + " is not a group");
    std::vector<std::string> out;
    for (auto& kv : ch) out.push_back(kv.first);
    return out;
  }
  H5Value read(const std::string& path) const {
    const uint64_t oh = resolve(path);
    H5Value v;
    bool have_space = false, have_type = false, have_layout = false;
    int layout_kind = 0;
    uint64_t addr = 0, lsize = 0;
    bool vlen = false;
    for (const Msg& m : messages(oh)) {
      if (m.type == 0x01) { v.shape = dataspace(m.data); have_space = true; }
      else if (m.type == 0x03) { datatype(m.data, v, vlen); have_type = true; }
      else if (m.type == 0x08) { layout(m.data, layout_kind, addr, lsize); have_layout = true; }
      else if (m.type == 0x0B) throw std::runtime_error("hdf5: filtered (compressed) dataset: " + path);
    }
    if (!have_space || !have_type || !have_layout) throw std::runtime_error("hdf5: " + path + " is not a dataset");
    const size_t n = v.count(), nbytes = n * (vlen ? 16 : (size_t)v.size);
    const unsigned char* src = nullptr;
    std::vector<unsigned char> zeros;
    if (layout_kind == 0) src = buf_.data() + addr;  // compact: addr is an absolute buffer offset
    else if (addr == UNDEF) { zeros.assign(nbytes, 0); src = zeros.data(); }
    else src = at(addr + base_, nbytes);
    fill(v, src, n, vlen);
    return v;
  }
  // attributes of a group or dataset
  std::map<std::string, H5Value> attrs(const std::string& path) const {
    std::map<std::string, H5Value> out;
    for (const Msg& m : messages(resolve(path))) {
      if (m.type != 0x0C) continue;
      size_t p = m.data;
      const int ver = u8(p);
      const size_t nsz = u16(p + 2), dsz = u16(p + 4), ssz = u16(p + 6);
      size_t q = p + 8 + (ver == 3 ? 1 : 0);
      if (ver < 1 || ver > 3) throw std::runtime_error("hdf5: attribute message version");
      auto pad = [&](size_t n) { return ver == 1 ? (n + 7) / 8 * 8 : n; };
      std::string name((const char*)at(q, nsz), strnlen((const char*)at(q, nsz), nsz));
      q += pad(nsz);
      H5Value v;
      bool vlen = false;
      datatype(q, v, vlen);
      q += pad(dsz);
      if (ssz >= 4) v.shape = dataspace(q);
      q += pad(ssz);
      const size_t n = v.count();
      fill(v, at(q, n * (vlen ? 16 : (size_t)v.size)), n, vlen);
      out[name] = v;
    }
    return out;
  }
  std::string attr_string(const std::string& path, const std::string& name) const {
    auto a = attrs(path);
    auto it = a.find(name);
    if (it == a.end() || it->second.cls != 3 || it->second.strings.empty()) return "";
    return it->second.strings[0];
  }

 private:
  static constexpr uint64_t UNDEF = ~uint64_t(0);
  struct Msg { int type; size_t data; size_t size; };
  std::vector<unsigned char> buf_;
  uint64_t base_ = 0, root_ = 0;

  const unsigned char* at(uint64_t p, size_t n) const {
    if (p > buf_.size() || n > buf_.size() - p) throw std::runtime_error("hdf5: read past the end of the file");
    return buf_.data() + p;
  }
  int u8(uint64_t p) const { return *at(p, 1); }
  uint32_t u16(uint64_t p) const { uint16_t v; std::memcpy(&v, at(p, 2), 2); return v; }
  uint32_t u32(uint64_t p) const { uint32_t v; std::memcpy(&v, at(p, 4), 4); return v; }
  uint64_t u64(uint64_t p) const { uint64_t v; std::memcpy(&v, at(p, 8), 8); return v; }

  std::vector<Msg> messages(uint64_t addr) const {
    if (u8(addr) != 1) throw std::runtime_error("hdf5: object header version " + std::to_string(u8(addr)) + " is not supported");
    const size_t nmsg = u16(addr + 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, u32(addr + 8)}};
    std::vector<Msg> out;
    for (size_t b = 0; b < blocks.size() && out.size() < nmsg; ++b) {
      uint64_t p = blocks[b].first;
      const uint64_t end = p + blocks[b].second;
      while (p + 8 <= end && out.size() < nmsg) {
        const int type = (int)u16(p);
        const size_t sz = u16(p + 2);
        if (type == 0x10) blocks.push_back({u64(p + 8) + base_, u64(p + 16)});
        out.push_back({type, (size_t)(p + 8), sz});
        p += 8 + sz;
      }
    }
    return out;
  }
  bool children(uint64_t oh, std::map<std::string, uint64_t>& out) const {
    uint64_t btree = UNDEF, heap = UNDEF;
    for (const Msg& m : messages(oh))
      if (m.type == 0x11) { btree = u64(m.data) + base_; heap = u64(m.data + 8) + base_; }
    if (btree == UNDEF) return false;
    if (std::memcmp(at(heap, 4), "HEAP", 4) != 0) throw std::runtime_error("hdf5: bad local heap");
    walk(btree, u64(heap + 24) + base_, out);
    return true;
  }
  void walk(uint64_t addr, uint64_t heap_data, std::map<std::string, uint64_t>& out) const {
    if (std::memcmp(at(addr, 4), "TREE", 4) != 0) throw std::runtime_error("hdf5: bad B-tree node");
    const int level = u8(addr + 5);
    const size_t n = u16(addr + 6);
    uint64_t p = addr + 24;
    for (size_t i = 0; i < n; ++i, p += 16) {
      const uint64_t child = u64(p + 8) + base_;
      if (level > 0) { walk(child, heap_data, out); continue; }
      if (std::memcmp(at(child, 4), "SNOD", 4) != 0) throw std::runtime_error("hdf5: bad symbol-table node");
      const size_t ns = u16(child + 6);
      for (size_t s = 0; s < ns; ++s) {
        const uint64_t e = child + 8 + 40 * s;
        const char* name = (const char*)at(heap_data + u64(e), 1);
        out[std::string(name)] = u64(e + 8) + base_;
      }
    }
  }
  uint64_t resolve(const std::string& path) const {
    uint64_t oh = root_ + base_;
    size_t i = 0;
    while (i < path.size()) {
      while (i < path.size() && path[i] == '/') ++i;
      size_t j = path.find('/', i);
      if (j == std::string::npos) j = path.size();
      if (j == i) break;
      std::map<std::string, uint64_t> ch;
      if (!children(oh, ch)) throw std::out_of_range("hdf5: no such object " + path);
      auto it = ch.find(path.substr(i, j - i));
      if (it == ch.end()) throw std::out_of_range("hdf5: no such object " + path);
      oh = it->second;
      i = j;
    }
    return oh;
  }
  std::vector<uint64_t> dataspace(uint64_t p) const {
    const int ver = u8(p), rank = u8(p + 1);
    if (ver != 1 && ver != 2) throw std::runtime_error("hdf5: dataspace version");
    const uint64_t q = p + (ver == 1 ? 8 : 4);
    std::vector<uint64_t> s(rank);
    for (int i = 0; i < rank; ++i) s[i] = u64(q + 8 * i);
    return s;
  }
  void datatype(uint64_t p, H5Value& v, bool& vlen) const {
    const int cls = u8(p) & 0x0F, bits0 = u8(p + 1);
    v.size = (int)u32(p + 4);
    vlen = false;
    if (cls == 0) { if (bits0 & 1) throw std::runtime_error("hdf5: big-endian data"); v.cls = 0; v.is_signed = (bits0 & 0x08) != 0; }
    else if (cls == 1) { if (bits0 & 1) throw std::runtime_error("hdf5: big-endian data"); v.cls = 1; }
    else if (cls == 3) v.cls = 3;
    else if (cls == 9 && (bits0 & 0x0F) == 1) { v.cls = 3; vlen = true; }
    else if (cls == 8) {  // enumeration (h5py booleans): read through its base integer type
      bool dummy;
      H5Value base;
      datatype(p + 8, base, dummy);
      v.cls = 0; v.is_signed = base.is_signed;
    } else throw std::runtime_error("hdf5: datatype class " + std::to_string(cls) + " is not supported");
  }
  void layout(uint64_t p, int& kind, uint64_t& addr, uint64_t& size) const {
    const int ver = u8(p);
    if (ver == 3) {
      const int cls = u8(p + 1);
      if (cls == 1) { kind = 1; addr = u64(p + 2); size = u64(p + 10); return; }
      if (cls == 0) { kind = 0; size = u16(p + 2); addr = p + 4; return; }
      throw std::runtime_error("hdf5: chunked layout is not supported");
    }
    if (ver == 1 || ver == 2) {
      if (u8(p + 2) != 1) throw std::runtime_error("hdf5: layout v1/v2 non-contiguous");
      kind = 1; addr = u64(p + 8); size = 0; return;
    }
    throw std::runtime_error("hdf5: layout version");
  }
  void fill(H5Value& v, const unsigned char* src, size_t n, bool vlen) const {
    if (v.cls != 3) { v.raw.assign(src, src + n * v.size); return; }
    v.strings.resize(n);
    for (size_t i = 0; i < n; ++i) {
      if (!vlen) {
        const char* s = (const char*)src + i * v.size;
        v.strings[i] = std::string(s, strnlen(s, v.size));
      } else {
        uint32_t len, idx;
        uint64_t gaddr;
        std::memcpy(&len, src + 16 * i, 4);
        std::memcpy(&gaddr, src + 16 * i + 4, 8);
        std::memcpy(&idx, src + 16 * i + 12, 4);
        v.strings[i] = gheap(gaddr + base_, idx).substr(0, len);
      }
    }
  }
  std::string gheap(uint64_t a, uint32_t idx) const {
    if (std::memcmp(at(a, 4), "GCOL", 4) != 0) throw std::runtime_error("hdf5: bad global heap");
    const uint64_t end = a + u64(a + 8);
    uint64_t p = a + 16;
    while (p + 16 <= end) {
      const uint32_t oid = u16(p);
      const uint64_t osz = u64(p + 8);
      if (oid == 0) break;
      if (oid == idx) return std::string((const char*)at(p + 16, osz), osz);
      p += 16 + (osz + 7) / 8 * 8;
    }
    throw std::runtime_error("hdf5: global heap object not found");
  }
};

}  // namespace qeio
