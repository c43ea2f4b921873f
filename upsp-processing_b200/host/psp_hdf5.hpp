// psp_hdf5.hpp -- the HDF5 outputs of psp_process (`-h5_out`, `extras.h5`) without an HDF5 library.
//
// Reference: upsp::PSPWriter cpp/lib/PSPHDF5.ipp:14-78 (file + root attributes psph5_version / nodal / transpose),
// :104-312 (write_grid: /Grid/x, y, z, triangles + components or grid_sizes, attribute units; root attribute
// structured), :361-411 (write_new_dataset + add_units), :447-551 (write_camera_settings), :555-751
// (write_tunnel_conditions + code_version), :755-781 (write_string_attribute); their use in
// cpp/exec/psp_process.cpp:2400-2420 and :2535-2604 (rms, average, coverage, steady_state, model_temp; the "frames"
// time histories are no longer written there: lines 1881-1923 and 2510-2511 are commented out in the reference, the flat
// files intensity_transpose / pressure_transpose carry them).
//
// No HDF5 library exists in this image, so the subset of the file format those calls produce is written by hand,
// following the HDF5 File Format Specification 2.0 and the byte layout of the reference's own fixtures
// (cpp/test/inputs/unstruct_nodal_pencil*.h5, HDF5 1.10 defaults): superblock version 0, old-style groups (one version-1
// B-tree node + local heap + ONE symbol-table node per group: the superblock's group leaf K is raised to 32 so that 64
// links fit), version-1 object headers, contiguous little-endian datasets of IEEE float / 32-bit integers /
// fixed-length strings with a version-2 fill-value message, version-1 attributes.  tests/h5min.py (a reader that parses
// the reference's fixtures) reads these files back in tests/test_psp_hdf5.py.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace upsp_b200 {

constexpr unsigned PSPH5_VERSION = 1;        // cpp/include/PSPHDF5.h
constexpr size_t H5_STRING_LEN = 256;

class H5Builder {
 public:
  enum Type { F32, F64, I32, U32, U16, STR };

  struct Attr {
    std::string name;
    Type type;
    std::vector<uint8_t> data;      // one element (a 256-byte string or one number)
  };
  struct Node {
    std::string name;
    bool group = false;
    std::vector<Attr> attrs;
    std::vector<Node> children;     // groups
    Type type = F32;                // datasets
    std::vector<uint64_t> dims;
    std::vector<uint8_t> data;
    bool fill = true;               // fill value 0 defined (the reference sets one on its numeric datasets)
  };

  Node root;
  H5Builder() { root.group = true; }

  static Attr attr_num(const std::string& name, Type t, uint64_t v) {
    Attr a{name, t, {}};
    const size_t n = size_of(t);
    a.data.resize(n);
    std::memcpy(a.data.data(), &v, n);        // little-endian host
    return a;
  }
  static Attr attr_str(const std::string& name, const std::string& v) {
    Attr a{name, STR, std::vector<uint8_t>(H5_STRING_LEN, 0)};
    std::memcpy(a.data.data(), v.data(), std::min(v.size(), H5_STRING_LEN));
    return a;
  }
  static size_t size_of(Type t) {
    switch (t) {
      case F32: case I32: case U32: return 4;
      case F64: return 8;
      case U16: return 2;
      case STR: return H5_STRING_LEN;
    }
    return 0;
  }

  Node& group(const std::string& name) {        // first-level groups only ("/Grid", "/Condition")
    for (auto& c : root.children)
      if (c.group && c.name == name) return c;
    Node g;
    g.name = name;
    g.group = true;
    root.children.push_back(std::move(g));
    return root.children.back();
  }
  template <typename T>
  static Node dataset(const std::string& name, Type t, std::vector<uint64_t> dims, const T* values, bool fill = true) {
    Node d;
    d.name = name;
    d.type = t;
    d.dims = std::move(dims);
    d.fill = fill;
    size_t n = 1;
    for (auto v : d.dims) n *= v;
    d.data.resize(n * size_of(t));
    if (n) std::memcpy(d.data.data(), values, d.data.size());
    return d;
  }
  static Node string_dataset(const std::string& name, const std::string& v) {
    std::vector<uint8_t> buf(H5_STRING_LEN, 0);
    std::memcpy(buf.data(), v.data(), std::min(v.size(), H5_STRING_LEN));
    return dataset<uint8_t>(name, STR, {1}, buf.data(), false);
  }

  void write(const std::string& path) {
    buf_.assign(96, 0);                            // superblock + root symbol-table entry
    const Placed r = place(root);
    // ---- superblock, version 0
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    std::memcpy(&buf_[0], sig, 8);
    buf_[13] = 8;                                  // size of offsets
    buf_[14] = 8;                                  // size of lengths
    put16(16, LEAF_K);
    put16(18, INT_K);
    put64(24, 0);                                  // base address
    put64(32, UNDEF);                              // free-space info
    put64(40, buf_.size());                        // end of file
    put64(48, UNDEF);                              // driver info
    put64(56, 0);                                  // root entry: link name offset
    put64(64, r.ohdr);
    put32(72, 1);                                  // cache type 1: B-tree + heap addresses in the scratch pad
    put64(80, r.btree);
    put64(88, r.heap);
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::invalid_argument("Cannot open " + path + " for writing");
    f.write(reinterpret_cast<const char*>(buf_.data()), (std::streamsize)buf_.size());
    if (!f) throw std::invalid_argument("Could not write " + path);
  }

 private:
  static constexpr uint64_t UNDEF = ~0ull;
  static constexpr unsigned LEAF_K = 32, INT_K = 16;
  struct Placed {
    uint64_t ohdr = 0, btree = 0, heap = 0;
  };
  std::vector<uint8_t> buf_;

  uint64_t alloc(size_t n) {
    const uint64_t at = (buf_.size() + 7) & ~(uint64_t)7;
    buf_.resize(at + n, 0);
    return at;
  }
  void put16(uint64_t at, uint16_t v) { std::memcpy(&buf_[at], &v, 2); }
  void put32(uint64_t at, uint32_t v) { std::memcpy(&buf_[at], &v, 4); }
  void put64(uint64_t at, uint64_t v) { std::memcpy(&buf_[at], &v, 8); }

  static void app(std::vector<uint8_t>& m, const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    m.insert(m.end(), b, b + n);
  }
  template <typename T>
  static void appv(std::vector<uint8_t>& m, T v) { app(m, &v, sizeof v); }
  static void pad8(std::vector<uint8_t>& m) { m.resize((m.size() + 7) & ~(size_t)7, 0); }

  static std::vector<uint8_t> msg_datatype(Type t) {
    std::vector<uint8_t> m;
    switch (t) {
      case F32: case F64: {
        const bool d = t == F64;
        const uint8_t head[4] = {0x11, 0x20, (uint8_t)(d ? 63 : 31), 0x00};       // class 1 v1; LE, msb-set norm.; sign bit
        app(m, head, 4);
        appv<uint32_t>(m, d ? 8 : 4);
        appv<uint16_t>(m, 0);                       // bit offset
        appv<uint16_t>(m, d ? 64 : 32);             // precision
        appv<uint8_t>(m, d ? 52 : 23);              // exponent location
        appv<uint8_t>(m, d ? 11 : 8);               // exponent size
        appv<uint8_t>(m, 0);                        // mantissa location
        appv<uint8_t>(m, d ? 52 : 23);              // mantissa size
        appv<uint32_t>(m, d ? 1023 : 127);          // bias
        break;
      }
      case I32: case U32: case U16: {
        const uint8_t head[4] = {0x10, (uint8_t)(t == I32 ? 0x08 : 0x00), 0x00, 0x00};   // class 0 v1; LE; signed bit 3
        app(m, head, 4);
        appv<uint32_t>(m, (uint32_t)size_of(t));
        appv<uint16_t>(m, 0);
        appv<uint16_t>(m, (uint16_t)(8 * size_of(t)));
        break;
      }
      case STR: {
        const uint8_t head[4] = {0x13, 0x00, 0x00, 0x00};                            // class 3 v1; null-terminated ASCII
        app(m, head, 4);
        appv<uint32_t>(m, (uint32_t)H5_STRING_LEN);
        break;
      }
    }
    return m;
  }
  static std::vector<uint8_t> msg_dataspace(const std::vector<uint64_t>& dims) {
    std::vector<uint8_t> m = {1, (uint8_t)dims.size(), 1, 0, 0, 0, 0, 0};     // version 1, rank, max dims present
    for (auto d : dims) appv<uint64_t>(m, d);
    for (auto d : dims) appv<uint64_t>(m, d);
    return m;
  }
  static void add_msg(std::vector<uint8_t>& hdr, unsigned& nmsg, uint16_t type, std::vector<uint8_t> body, uint8_t flags = 0) {
    pad8(body);
    appv<uint16_t>(hdr, type);
    appv<uint16_t>(hdr, (uint16_t)body.size());
    appv<uint8_t>(hdr, flags);
    const uint8_t rsv[3] = {0, 0, 0};
    app(hdr, rsv, 3);
    app(hdr, body.data(), body.size());
    ++nmsg;
  }
  static void add_attrs(std::vector<uint8_t>& hdr, unsigned& nmsg, const std::vector<Attr>& attrs) {
    for (const Attr& a : attrs) {
      std::vector<uint8_t> dt = msg_datatype(a.type), ds = msg_dataspace({1});
      std::vector<uint8_t> m;
      appv<uint8_t>(m, 1);                                    // version 1
      appv<uint8_t>(m, 0);
      appv<uint16_t>(m, (uint16_t)(a.name.size() + 1));
      appv<uint16_t>(m, (uint16_t)dt.size());
      appv<uint16_t>(m, (uint16_t)ds.size());
      app(m, a.name.c_str(), a.name.size() + 1);
      pad8(m);
      app(m, dt.data(), dt.size());
      pad8(m);
      app(m, ds.data(), ds.size());
      pad8(m);
      app(m, a.data.data(), a.data.size());
      if (m.size() > 65000) throw std::invalid_argument("attribute " + a.name + " too large");
      add_msg(hdr, nmsg, 0x000C, std::move(m));
    }
  }
  uint64_t emit_header(const std::vector<uint8_t>& msgs, unsigned nmsg) {
    const uint64_t at = alloc(16 + msgs.size());
    buf_[at] = 1;                                             // version 1
    put16(at + 2, (uint16_t)nmsg);
    put32(at + 4, 1);                                         // reference count
    put32(at + 8, (uint32_t)msgs.size());
    std::memcpy(&buf_[at + 16], msgs.data(), msgs.size());
    return at;
  }

  Placed place(const Node& n) {
    Placed out;
    std::vector<uint8_t> msgs;
    unsigned nmsg = 0;
    if (!n.group) {
      const uint64_t daddr = n.data.empty() ? UNDEF : alloc(n.data.size());
      if (!n.data.empty()) std::memcpy(&buf_[daddr], n.data.data(), n.data.size());
      add_msg(msgs, nmsg, 0x0001, msg_dataspace(n.dims));
      add_msg(msgs, nmsg, 0x0003, msg_datatype(n.type), 1);                  // constant
      {
        std::vector<uint8_t> f = {2, 2, 2, 1};                                // version 2, late alloc, write if set, defined
        const uint32_t fs = (n.fill && n.type != STR) ? (uint32_t)size_of(n.type) : 0;
        appv<uint32_t>(f, fs);
        f.resize(f.size() + fs, 0);                                            // fill value 0
        add_msg(msgs, nmsg, 0x0005, std::move(f), 1);
      }
      {
        std::vector<uint8_t> l = {3, 1};                                      // version 3, contiguous
        appv<uint64_t>(l, daddr);
        appv<uint64_t>(l, (uint64_t)n.data.size());
        add_msg(msgs, nmsg, 0x0008, std::move(l));
      }
      add_attrs(msgs, nmsg, n.attrs);
      out.ohdr = emit_header(msgs, nmsg);
      return out;
    }
    // ---- group: children first, then local heap, symbol-table node, B-tree node, object header
    std::vector<std::pair<std::string, Placed>> kids;
    std::vector<const Node*> order;
    for (const Node& c : n.children) order.push_back(&c);
    std::sort(order.begin(), order.end(), [](const Node* a, const Node* b) { return a->name < b->name; });
    if (order.size() > 2 * LEAF_K) throw std::invalid_argument("group " + n.name + ": too many links for one symbol-table node");
    for (const Node* c : order) kids.emplace_back(c->name, place(*c));
    std::vector<uint8_t> names(8, 0);                         // offset 0: the empty string
    std::vector<uint64_t> noff;
    for (auto& k : kids) {
      noff.push_back(names.size());
      app(names, k.first.c_str(), k.first.size() + 1);
      pad8(names);
    }
    const uint64_t dseg = alloc(names.size());
    std::memcpy(&buf_[dseg], names.data(), names.size());
    out.heap = alloc(32);
    std::memcpy(&buf_[out.heap], "HEAP", 4);
    put64(out.heap + 8, names.size());
    put64(out.heap + 16, 1);                                  // H5HL_FREE_NULL: no free block
    put64(out.heap + 24, dseg);
    const uint64_t snod = alloc(8 + 40 * 2 * LEAF_K);
    std::memcpy(&buf_[snod], "SNOD", 4);
    buf_[snod + 4] = 1;
    put16(snod + 6, (uint16_t)kids.size());
    for (size_t i = 0; i < kids.size(); ++i) {
      const uint64_t e = snod + 8 + 40 * i;
      put64(e, noff[i]);
      put64(e + 8, kids[i].second.ohdr);
      if (order[i]->group) {
        put32(e + 16, 1);
        put64(e + 24, kids[i].second.btree);
        put64(e + 32, kids[i].second.heap);
      }
    }
    out.btree = alloc(24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8);
    std::memcpy(&buf_[out.btree], "TREE", 4);
    buf_[out.btree + 4] = 0;                                  // group node
    buf_[out.btree + 5] = 0;                                  // level 0
    put16(out.btree + 6, kids.empty() ? 0 : 1);
    put64(out.btree + 8, UNDEF);
    put64(out.btree + 16, UNDEF);
    if (!kids.empty()) {
      put64(out.btree + 24, 0);                               // key 0: the empty string
      put64(out.btree + 32, snod);
      put64(out.btree + 40, noff.back());                     // key 1: the largest name of the node
    }
    {
      std::vector<uint8_t> st;
      appv<uint64_t>(st, out.btree);
      appv<uint64_t>(st, out.heap);
      add_msg(msgs, nmsg, 0x0011, std::move(st));
    }
    add_attrs(msgs, nmsg, n.attrs);
    out.ohdr = emit_header(msgs, nmsg);
    return out;
  }
};

// ---- the reference's writer interface on top of it -------------------------------------------------------------------
struct H5TunnelConditions {     // upsp::TunnelConditions, cpp/include/non_cv_upsp.h
  std::string test_id;
  int run = 0, seq = 0;
  float alpha = 0, beta = 0, phi = 0, mach = 0, rey = 0, ptot = 0, qbar = 0, ttot = 0, tcavg = 0, ps = 0;
};
struct H5CameraSettings {       // upsp::CameraSettings
  int framerate = 0;
  float fstop = 0, exposure = 0;
  std::vector<float> focal_lengths;
};

class PSPWriter {
 public:
  // nodal files only (psp_process writes nodal data); `transposed` is the attribute psp_process sets for -h5_out
  PSPWriter(std::string filename, size_t data_pts, bool transposed = false, bool nodal = true)
      : filename_(std::move(filename)), data_pts_(data_pts) {
    h5_.root.attrs.push_back(H5Builder::attr_num("psph5_version", H5Builder::U32, PSPH5_VERSION));
    h5_.root.attrs.push_back(H5Builder::attr_num("nodal", H5Builder::U16, nodal ? 1 : 0));
    h5_.root.attrs.push_back(H5Builder::attr_num("transpose", H5Builder::U16, transposed ? 1 : 0));
    h5_.group("Condition");
  }
  void write_unstructured_grid(const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& z,
                               const std::vector<unsigned>& tri_nodes, const std::vector<int>& comps, const std::string& units) {
    grid_common(0, x, y, z, units);
    auto& g = h5_.group("Grid");
    g.children.push_back(H5Builder::dataset("triangles", H5Builder::U32, {tri_nodes.size() / 3, 3}, tri_nodes.data(), false));
    g.children.push_back(H5Builder::dataset("components", H5Builder::I32, {comps.size()}, comps.data(), false));
  }
  void write_structured_grid(const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& z,
                             const std::vector<int>& zone_sizes /* [zones][3] */, const std::string& units) {
    grid_common(1, x, y, z, units);
    h5_.group("Grid").children.push_back(
        H5Builder::dataset("grid_sizes", H5Builder::I32, {zone_sizes.size() / 3, 3}, zone_sizes.data(), false));
  }
  void write_tunnel_conditions(const H5TunnelConditions& c) {
    auto& g = h5_.group("Condition");
    g.children.push_back(H5Builder::string_dataset("test_id", c.test_id));
    add_scalar(g, "run", c.run, "-");
    add_scalar(g, "sequence", c.seq, "-");
    add_scalar(g, "alpha", c.alpha, "deg");
    add_scalar(g, "beta", c.beta, "deg");
    add_scalar(g, "phi", c.phi, "deg");
    add_scalar(g, "mach", c.mach, "-");
    add_scalar(g, "reynolds_number", c.rey, "millions/ft");
    add_scalar(g, "total_pressure", c.ptot, "psf");
    add_scalar(g, "dynamic_pressure", c.qbar, "psf");
    add_scalar(g, "total_temperature", c.ttot, "degF");
    add_scalar(g, "thermocouple_average_temperature", c.tcavg, "degF");
    add_scalar(g, "static_pressure", c.ps, "psf");
  }
  void write_camera_settings(const H5CameraSettings& cs) {
    auto& g = h5_.group("Condition");
    add_scalar(g, "frame_rate", cs.framerate, "Hz");
    add_scalar(g, "fstop", cs.fstop, "-");
    add_scalar(g, "exposure", cs.exposure, "microseconds");
    auto d = H5Builder::dataset("focal_length", H5Builder::F32, {cs.focal_lengths.size()}, cs.focal_lengths.data());
    d.attrs.push_back(H5Builder::attr_str("units", "mm"));
    g.children.push_back(std::move(d));
  }
  void write_string_attribute(const std::string& name, const std::string& value) {
    h5_.root.attrs.push_back(H5Builder::attr_str(name, value));
  }
  void write_new_dataset(const std::string& name, const std::vector<float>& sol, const std::string& units = "") {
    if (sol.size() != data_pts_) throw std::invalid_argument("Must have solution at every grid point");
    auto d = H5Builder::dataset(name, H5Builder::F32, {sol.size()}, sol.data());
    if (!units.empty()) d.attrs.push_back(H5Builder::attr_str("units", units));
    h5_.root.children.push_back(std::move(d));
  }
  void close() { h5_.write(filename_); }

 private:
  std::string filename_;
  size_t data_pts_;
  H5Builder h5_;

  void grid_common(int structured, const std::vector<float>& x, const std::vector<float>& y, const std::vector<float>& z,
                   const std::string& units) {
    h5_.root.attrs.push_back(H5Builder::attr_num("structured", H5Builder::U16, (uint64_t)structured));
    auto& g = h5_.group("Grid");
    g.children.push_back(H5Builder::dataset("x", H5Builder::F32, {x.size()}, x.data()));
    g.children.push_back(H5Builder::dataset("y", H5Builder::F32, {y.size()}, y.data()));
    g.children.push_back(H5Builder::dataset("z", H5Builder::F32, {z.size()}, z.data()));
    g.attrs.push_back(H5Builder::attr_str("units", units));
  }
  static void add_scalar(H5Builder::Node& g, const std::string& name, int v, const std::string& units) {
    auto d = H5Builder::dataset(name, H5Builder::I32, {1}, &v);
    d.attrs.push_back(H5Builder::attr_str("units", units));
    g.children.push_back(std::move(d));
  }
  static void add_scalar(H5Builder::Node& g, const std::string& name, float v, const std::string& units) {
    auto d = H5Builder::dataset(name, H5Builder::F32, {1}, &v);
    d.attrs.push_back(H5Builder::attr_str("units", units));
    g.children.push_back(std::move(d));
  }
};

}  // namespace upsp_b200
