// h5_append.hpp -- add a chunked [rows x extent] float dataset to an existing HDF5 file, in place, without an HDF5
// library: what the reference's `add_field` does with H5::H5File(H5F_ACC_RDWR) + createDataSet + hyperslab writes.
//
// Reference: cpp/exec/add_field.cpp:22-128 (arguments, flat-file checks, creation property list: NATIVE_FLOAT, fill
// value 0.0, chunk {1, extent}, dataspace {nRows, extent}, rows written in file order); its use
// docs/sphinx/quick-start.rst:125-160 (`add_field <h5> frames <pressure_transpose> <number of frames>`).
//
// File format (HDF5 File Format Specification 2.0; the subset HDF5 1.10 writes with default settings, which is what
// PSPWriter files -- the reference's and host/psp_hdf5.hpp's -- are): superblock version 0 with 8-byte offsets / lengths,
// old-style groups (version-1 B-tree of symbol-table nodes + local heap), version-1 object headers.  The new dataset is
//   * raw chunks: row i of the flat file at `data + i * extent * 4` (unfiltered, appended at the end of the file),
//   * a version-1 chunk B-tree (node type 1, 2K = 64 entries per node, K = 32: the library's default for indexed storage,
//     which superblock version 0 cannot override), keys (chunk bytes, filter mask 0, offsets {row, 0, 0}), the last key of
//     the tree {rows, extent, 4} with size 0 as the library writes it (checked against the chunk trees of the reference's
//     fixtures cpp/test/inputs/unstruct_nodal_pencil*.h5),
//   * a version-1 object header: dataspace (version 1, max dims = dims), datatype (IEEE float LE), fill value (version 2,
//     incremental allocation, written if set, 0.0) and the old fill-value message the library adds beside it, layout
//     (version 3, chunked, dimensionality 3: {1, extent, 4}), modification time.
// The root group is re-linked rather than edited: every link of the root group is read (B-tree walk), the new name is
// added, and a fresh local heap + symbol-table nodes + B-tree are appended; then the root object header's symbol-table
// message, the superblock's cached copy of it and the end-of-file address are patched.  The old group structures stay in
// the file as unreferenced space (the library leaves such holes too).  Everything is appended before anything is patched.
//
// Verification: tests/test_add_field.py reads the result with tests/h5min.py (the reader held against the reference's
// fixtures): on files written by host/psp_hdf5.hpp and on copies of the reference's own fixtures (written by libhdf5).
// Not verified with libhdf5 itself (absent from this image).
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <vector>

namespace upsp_b200 {

class H5Appender {
 public:
  struct Link {
    std::string name;
    uint64_t ohdr = 0;
    uint32_t cache_type = 0;
    uint8_t scratch[16] = {0};
  };

  explicit H5Appender(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDWR);
    if (fd_ < 0) throw std::runtime_error("unable to open file");
    try {
      parse();
    } catch (...) {
      ::close(fd_);
      throw;
    }
  }
  ~H5Appender() {
    if (fd_ >= 0) ::close(fd_);
  }
  H5Appender(const H5Appender&) = delete;
  H5Appender& operator=(const H5Appender&) = delete;

  const std::vector<Link>& root_links() const { return links_; }
  bool has(const std::string& name) const {
    for (const Link& l : links_)
      if (l.name == name) return true;
    return false;
  }

  // Appends dataset `name` = the rows of `flat_fd` (rows x extent floats, file order) and links it into the root group.
  void add_chunked_float_dataset(const std::string& name_in, int flat_fd, uint64_t rows, uint64_t extent) {
    std::string name = name_in;
    while (!name.empty() && name[0] == '/') name.erase(0, 1);
    if (name.empty() || name.find('/') != std::string::npos)
      throw std::runtime_error("only datasets of the root group are supported: " + name_in);
    if (has(name)) throw std::runtime_error("unable to create dataset: name already exists");
    if (extent == 0 || extent * 4 > 0xffffffffull) throw std::runtime_error("chunk of " + std::to_string(extent) + " floats is not representable");
    const uint64_t row_bytes = extent * 4;
    cursor_ = align8(eof_);
    // ---- raw chunks
    const uint64_t data = cursor_;
    copy_from(flat_fd, rows * row_bytes);
    // ---- chunk B-tree
    const uint64_t btree = rows ? write_chunk_tree(data, rows, extent) : UNDEF;
    // ---- object header
    std::vector<uint8_t> msgs;
    unsigned nmsg = 0;
    {
      std::vector<uint8_t> m = {1, 2, 1, 0, 0, 0, 0, 0};        // dataspace version 1, rank 2, max dims present
      appv<uint64_t>(m, rows);
      appv<uint64_t>(m, extent);
      appv<uint64_t>(m, rows);
      appv<uint64_t>(m, extent);
      add_msg(msgs, nmsg, 0x0001, m);
    }
    {
      std::vector<uint8_t> m = {0x11, 0x20, 31, 0x00};            // class 1 version 1; little-endian, implied msb; sign bit 31
      appv<uint32_t>(m, 4);
      appv<uint16_t>(m, 0);
      appv<uint16_t>(m, 32);
      appv<uint8_t>(m, 23);
      appv<uint8_t>(m, 8);
      appv<uint8_t>(m, 0);
      appv<uint8_t>(m, 23);
      appv<uint32_t>(m, 127);
      add_msg(msgs, nmsg, 0x0003, m, 1);
    }
    {
      std::vector<uint8_t> m = {2, 3, 2, 1};                      // fill value version 2: incremental allocation, write if set, defined
      appv<uint32_t>(m, 4);
      appv<float>(m, 0.0f);
      add_msg(msgs, nmsg, 0x0005, m, 1);
    }
    {
      std::vector<uint8_t> m;                                     // old fill-value message, as the library writes beside the new one
      appv<uint32_t>(m, 4);
      appv<float>(m, 0.0f);
      add_msg(msgs, nmsg, 0x0004, m);
    }
    {
      std::vector<uint8_t> m = {3, 2, 3};                         // layout version 3, chunked, dimensionality rank + 1
      appv<uint64_t>(m, btree);
      appv<uint32_t>(m, 1);
      appv<uint32_t>(m, (uint32_t)extent);
      appv<uint32_t>(m, 4);
      add_msg(msgs, nmsg, 0x0008, m);
    }
    {
      std::vector<uint8_t> m = {1, 0, 0, 0};                      // modification time version 1
      appv<uint32_t>(m, (uint32_t)std::time(nullptr));
      add_msg(msgs, nmsg, 0x0012, m);
    }
    const uint64_t ohdr = emit_header(msgs, nmsg);
    // ---- the root group with the new link
    Link nl;
    nl.name = name;
    nl.ohdr = ohdr;
    std::vector<Link> all = links_;
    all.push_back(nl);
    std::sort(all.begin(), all.end(), [](const Link& a, const Link& b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    uint64_t new_btree = 0, new_heap = 0;
    write_group(all, new_btree, new_heap);
    if (::fsync(fd_) != 0) throw std::runtime_error("fsync failed");
    // ---- patch: symbol-table message of the root object header, superblock's cached copy, end-of-file address
    put64_at(symtab_msg_at_, new_btree);
    put64_at(symtab_msg_at_ + 8, new_heap);
    if (root_cache_type_ == 1) {
      put64_at(root_entry_at_ + 24, new_btree);
      put64_at(root_entry_at_ + 32, new_heap);
    }
    put64_at(eof_at_, cursor_);
    if (::fsync(fd_) != 0) throw std::runtime_error("fsync failed");
    eof_ = cursor_;
    links_ = all;
  }

 private:
  static constexpr uint64_t UNDEF = ~0ull;
  static constexpr unsigned CHUNK_K = 32;      // H5D default indexed-storage K (superblock version 0 has no field for it)
  std::string path_;
  int fd_ = -1;
  unsigned leaf_k_ = 4, int_k_ = 16;
  uint64_t base_ = 0, eof_ = 0, eof_at_ = 0, root_entry_at_ = 0, symtab_msg_at_ = 0, cursor_ = 0;
  uint32_t root_cache_type_ = 0;
  std::vector<Link> links_;

  static uint64_t align8(uint64_t v) { return (v + 7) & ~7ull; }
  template <typename T>
  static void appv(std::vector<uint8_t>& m, T v) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
    m.insert(m.end(), b, b + sizeof v);
  }
  static void add_msg(std::vector<uint8_t>& hdr, unsigned& nmsg, uint16_t type, std::vector<uint8_t> body, uint8_t flags = 0) {
    body.resize((body.size() + 7) & ~(size_t)7, 0);
    appv<uint16_t>(hdr, type);
    appv<uint16_t>(hdr, (uint16_t)body.size());
    appv<uint8_t>(hdr, flags);
    hdr.insert(hdr.end(), 3, 0);
    hdr.insert(hdr.end(), body.begin(), body.end());
    ++nmsg;
  }

  // ---- raw file access
  void read_at(uint64_t at, void* dst, size_t n) const {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
      const ssize_t k = ::pread(fd_, p, n, (off_t)(base_ + at));
      if (k <= 0) throw std::runtime_error("truncated file (read of " + std::to_string(n) + " bytes at " + std::to_string(at) + ")");
      p += k;
      at += (uint64_t)k;
      n -= (size_t)k;
    }
  }
  void write_at(uint64_t at, const void* src, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(src);
    while (n) {
      const ssize_t k = ::pwrite(fd_, p, n, (off_t)(base_ + at));
      if (k <= 0) throw std::runtime_error("write failed");
      p += k;
      at += (uint64_t)k;
      n -= (size_t)k;
    }
  }
  template <typename T>
  T get(uint64_t at) const {
    T v;
    read_at(at, &v, sizeof v);
    return v;
  }
  void put64_at(uint64_t at, uint64_t v) { write_at(at, &v, 8); }
  uint64_t append(const std::vector<uint8_t>& b) {
    const uint64_t at = align8(cursor_);
    if (at > cursor_) {
      const uint8_t z[8] = {0};
      write_at(cursor_, z, (size_t)(at - cursor_));
    }
    write_at(at, b.data(), b.size());
    cursor_ = at + b.size();
    return at;
  }
  void copy_from(int src_fd, uint64_t n) {
    std::vector<uint8_t> buf(8u << 20);
    uint64_t done = 0;
    while (done < n) {
      const size_t want = (size_t)std::min<uint64_t>(buf.size(), n - done);
      const ssize_t k = ::pread(src_fd, buf.data(), want, (off_t)done);
      if (k <= 0) throw std::runtime_error("flat file shorter than expected");
      write_at(cursor_, buf.data(), (size_t)k);
      cursor_ += (uint64_t)k;
      done += (uint64_t)k;
    }
  }
  uint64_t emit_header(const std::vector<uint8_t>& msgs, unsigned nmsg) {
    std::vector<uint8_t> h(16, 0);
    h[0] = 1;
    const uint16_t n16 = (uint16_t)nmsg;
    const uint32_t one = 1, sz = (uint32_t)msgs.size();
    std::memcpy(&h[2], &n16, 2);
    std::memcpy(&h[4], &one, 4);
    std::memcpy(&h[8], &sz, 4);
    h.insert(h.end(), msgs.begin(), msgs.end());
    return append(h);
  }

  // ---- reading the existing file
  void parse() {
    // the superblock may sit at 0, 512, 1024, ... (user block); PSPWriter files have none
    uint8_t sb[96];
    base_ = 0;
    read_at(0, sb, sizeof sb);
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (std::memcmp(sb, sig, 8) != 0) throw std::runtime_error("not an HDF5 file");
    if (sb[8] != 0) throw std::runtime_error("HDF5 superblock version " + std::to_string((int)sb[8]) + " is not supported (version 0 only)");
    if (sb[13] != 8 || sb[14] != 8) throw std::runtime_error("only 8-byte offsets and lengths are supported");
    uint16_t lk, ik;
    std::memcpy(&lk, sb + 16, 2);
    std::memcpy(&ik, sb + 18, 2);
    leaf_k_ = lk;
    int_k_ = ik;
    uint64_t base;
    std::memcpy(&base, sb + 24, 8);
    if (base != 0) throw std::runtime_error("non-zero base address is not supported");
    eof_at_ = 40;
    std::memcpy(&eof_, sb + 40, 8);
    struct stat st;
    if (::fstat(fd_, &st) != 0 || (uint64_t)st.st_size < eof_) throw std::runtime_error("file is shorter than its end-of-file address");
    root_entry_at_ = 56;
    uint64_t root_ohdr;
    std::memcpy(&root_ohdr, sb + 64, 8);
    std::memcpy(&root_cache_type_, sb + 72, 4);
    // root object header: find the symbol-table message (and where it sits)
    uint64_t btree = UNDEF, heap = UNDEF;
    {
      const uint8_t ver = get<uint8_t>(root_ohdr);
      if (ver != 1) throw std::runtime_error("root object header version " + std::to_string((int)ver) + " is not supported");
      const uint16_t nmsg = get<uint16_t>(root_ohdr + 2);
      const uint32_t hsize = get<uint32_t>(root_ohdr + 8);
      std::vector<std::pair<uint64_t, uint64_t>> blocks = {{root_ohdr + 16, hsize}};
      unsigned seen = 0;
      for (size_t bi = 0; bi < blocks.size() && seen < nmsg; ++bi) {
        uint64_t pos = blocks[bi].first;
        const uint64_t end = pos + blocks[bi].second;
        while (pos + 8 <= end && seen < nmsg) {
          const uint16_t type = get<uint16_t>(pos), size = get<uint16_t>(pos + 2);
          if (type == 0x0010) blocks.emplace_back(get<uint64_t>(pos + 8), get<uint64_t>(pos + 16));
          if (type == 0x0011) {
            symtab_msg_at_ = pos + 8;
            btree = get<uint64_t>(pos + 8);
            heap = get<uint64_t>(pos + 16);
          }
          pos += 8 + size;
          ++seen;
        }
      }
    }
    if (symtab_msg_at_ == 0) throw std::runtime_error("the root group is not an old-style group (no symbol-table message)");
    // local heap
    char hs[4];
    read_at(heap, hs, 4);
    if (std::memcmp(hs, "HEAP", 4) != 0) throw std::runtime_error("local heap signature");
    const uint64_t dseg_size = get<uint64_t>(heap + 8), dseg = get<uint64_t>(heap + 24);
    std::vector<char> names((size_t)dseg_size + 1, 0);
    read_at(dseg, names.data(), (size_t)dseg_size);
    walk_group(btree, names);
  }
  void walk_group(uint64_t addr, const std::vector<char>& names) {
    char sg[4];
    read_at(addr, sg, 4);
    if (std::memcmp(sg, "SNOD", 4) == 0) {
      const uint16_t n = get<uint16_t>(addr + 6);
      for (unsigned i = 0; i < n; ++i) {
        const uint64_t e = addr + 8 + 40ull * i;
        Link l;
        const uint64_t noff = get<uint64_t>(e);
        if (noff >= names.size()) throw std::runtime_error("link name outside the local heap");
        l.name = std::string(&names[(size_t)noff]);
        l.ohdr = get<uint64_t>(e + 8);
        l.cache_type = get<uint32_t>(e + 16);
        read_at(e + 24, l.scratch, 16);
        links_.push_back(l);
      }
      return;
    }
    if (std::memcmp(sg, "TREE", 4) != 0) throw std::runtime_error("group B-tree signature");
    if (get<uint8_t>(addr + 4) != 0) throw std::runtime_error("group B-tree node type");
    const uint16_t used = get<uint16_t>(addr + 6);
    for (unsigned i = 0; i < used; ++i) walk_group(get<uint64_t>(addr + 24 + 8 + 16ull * i), names);
  }

  // ---- chunk B-tree (node type 1): bottom-up, nodes of up to 2K entries, siblings linked per level
  uint64_t write_chunk_tree(uint64_t data, uint64_t rows, uint64_t extent) {
    struct Key {
      uint32_t nbytes;
      uint64_t off[3];
    };
    struct Ent {
      Key key;
      uint64_t child;
    };
    const uint32_t row_bytes = (uint32_t)(extent * 4);
    const Key last = {0, {rows, extent, 4}};
    const size_t cap = 2 * CHUNK_K, key_size = 8 + 8 * 3, node_size = 24 + (cap + 1) * key_size + cap * 8;
    std::vector<Ent> level;
    level.reserve((size_t)rows);
    for (uint64_t i = 0; i < rows; ++i) level.push_back({{row_bytes, {i, 0, 0}}, data + i * row_bytes});
    for (unsigned depth = 0;; ++depth) {
      const size_t n = level.size(), nnodes = (n + cap - 1) / cap, per = (n + nnodes - 1) / nnodes;
      const uint64_t first = align8(cursor_);
      std::vector<Ent> up;
      std::vector<uint8_t> node;
      for (size_t j = 0, at = 0; j < nnodes; ++j) {
        const size_t cnt = std::min(per, n - at);
        node.assign(node_size, 0);
        std::memcpy(&node[0], "TREE", 4);
        node[4] = 1;
        node[5] = (uint8_t)depth;
        const uint16_t used = (uint16_t)cnt;
        std::memcpy(&node[6], &used, 2);
        const uint64_t left = j ? first + (j - 1) * node_size : UNDEF, right = j + 1 < nnodes ? first + (j + 1) * node_size : UNDEF;
        std::memcpy(&node[8], &left, 8);
        std::memcpy(&node[16], &right, 8);
        size_t pos = 24;
        auto put_key = [&](const Key& k) {
          const uint32_t mask = 0;
          std::memcpy(&node[pos], &k.nbytes, 4);
          std::memcpy(&node[pos + 4], &mask, 4);
          std::memcpy(&node[pos + 8], k.off, 24);
          pos += key_size;
        };
        for (size_t i = 0; i < cnt; ++i) {
          put_key(level[at + i].key);
          std::memcpy(&node[pos], &level[at + i].child, 8);
          pos += 8;
        }
        put_key(at + cnt < n ? level[at + cnt].key : last);      // right key: the next node's first key, or the end of the dataset
        const uint64_t addr = append(node);
        if (addr != first + j * node_size) throw std::runtime_error("internal: chunk B-tree node placement");
        up.push_back({level[at].key, addr});
        at += cnt;
      }
      if (nnodes == 1) return first;
      level.swap(up);
    }
  }

  // ---- group: local heap, symbol-table nodes, B-tree (node type 0)
  void write_group(const std::vector<Link>& all, uint64_t& btree_out, uint64_t& heap_out) {
    // heap data segment: offset 0 = the empty string, names 8-byte aligned
    std::vector<uint8_t> seg(8, 0);
    std::vector<uint64_t> noff;
    for (const Link& l : all) {
      noff.push_back(seg.size());
      seg.insert(seg.end(), l.name.begin(), l.name.end());
      seg.push_back(0);
      seg.resize((seg.size() + 7) & ~(size_t)7, 0);
    }
    const uint64_t dseg = append(seg);
    std::vector<uint8_t> hp(32, 0);
    std::memcpy(&hp[0], "HEAP", 4);
    const uint64_t seg_size = seg.size(), free_null = 1;
    std::memcpy(&hp[8], &seg_size, 8);
    std::memcpy(&hp[16], &free_null, 8);         // H5HL_FREE_NULL: no free block in the data segment
    std::memcpy(&hp[24], &dseg, 8);
    heap_out = append(hp);
    // symbol-table nodes: as few as fit, evenly filled (every node of a split tree holds at least K entries)
    struct Ent {
      uint64_t left_key, right_key, child;     // heap offsets: largest name left of the child / largest name inside it
    };
    const size_t n = all.size(), cap = 2 * leaf_k_, nsnod = std::max<size_t>(1, (n + cap - 1) / cap), per = (n + nsnod - 1) / nsnod;
    std::vector<Ent> level;
    for (size_t j = 0, at = 0; j < nsnod; ++j) {
      const size_t cnt = std::min(per, n - at);
      std::vector<uint8_t> sn(8 + 40 * cap, 0);
      std::memcpy(&sn[0], "SNOD", 4);
      sn[4] = 1;
      const uint16_t c16 = (uint16_t)cnt;
      std::memcpy(&sn[6], &c16, 2);
      for (size_t i = 0; i < cnt; ++i) {
        uint8_t* e = &sn[8 + 40 * i];
        const Link& l = all[at + i];
        std::memcpy(e, &noff[at + i], 8);
        std::memcpy(e + 8, &l.ohdr, 8);
        std::memcpy(e + 16, &l.cache_type, 4);
        std::memcpy(e + 24, l.scratch, 16);
      }
      const uint64_t addr = append(sn);
      level.push_back({at ? noff[at - 1] : 0, cnt ? noff[at + cnt - 1] : 0, addr});
      at += cnt;
    }
    // B-tree levels
    const size_t bcap = 2 * int_k_, node_size = 24 + (bcap + 1) * 8 + bcap * 8;
    for (unsigned depth = 0;; ++depth) {
      const size_t m = level.size(), nnodes = (m + bcap - 1) / bcap, per_node = (m + nnodes - 1) / nnodes;
      const uint64_t first = align8(cursor_);
      std::vector<Ent> up;
      for (size_t j = 0, at = 0; j < nnodes; ++j) {
        const size_t cnt = std::min(per_node, m - at);
        std::vector<uint8_t> node(node_size, 0);
        std::memcpy(&node[0], "TREE", 4);
        node[4] = 0;
        node[5] = (uint8_t)depth;
        const uint16_t used = (uint16_t)(n ? cnt : 0);
        std::memcpy(&node[6], &used, 2);
        const uint64_t left = j ? first + (j - 1) * node_size : UNDEF, right = j + 1 < nnodes ? first + (j + 1) * node_size : UNDEF;
        std::memcpy(&node[8], &left, 8);
        std::memcpy(&node[16], &right, 8);
        size_t pos = 24;
        if (n) {
          std::memcpy(&node[pos], &level[at].left_key, 8);
          pos += 8;
          for (size_t i = 0; i < cnt; ++i) {
            std::memcpy(&node[pos], &level[at + i].child, 8);
            std::memcpy(&node[pos + 8], &level[at + i].right_key, 8);
            pos += 16;
          }
        }
        const uint64_t addr = append(node);
        if (addr != first + j * node_size) throw std::runtime_error("internal: group B-tree node placement");
        up.push_back({level[at].left_key, level[at + cnt - 1].right_key, addr});
        at += cnt;
      }
      if (nnodes == 1) {
        btree_out = first;
        return;
      }
      level.swap(up);
    }
  }
};

}  // namespace upsp_b200
