// Self-contained reader / writer for the HDF5 subset GauXC's records use (no libhdf5 in this image):
// superblock v0, v1 object headers (+ continuation blocks), symbol-table groups (v1 B-tree, local heap,
// SNOD), dataspace v1/v2, datatype classes 0 (fixed), 1 (float), 6 (compound v1-v3), 10 (array v2/v3),
// data layout v3 contiguous / compact.  Replaces src/external/hdf5_read.cxx:47-156 / hdf5_write.cxx:22-86 of
// the reference (HighFive + libhdf5); the byte layout of what the writer emits mirrors the reference's own
// fixture files (tests/ref_data/*.hdf5), compound member names and offsets included (hdf5_util.hpp:27-62).
#include "hdf5_io.hpp"
#include <algorithm>
#include <cstring>
#include <fstream>
#include <map>

namespace GauXC {

namespace {

constexpr uint64_t UNDEF = ~0ull;

struct Reader {
  std::vector<uint8_t> b;
  uint64_t root_ohdr = 0;

  template <typename T>
  T get(uint64_t off) const {
    if (off + sizeof(T) > b.size()) GAUXC_GENERIC_EXCEPTION("HDF5: read past the end of the file");
    T v;
    std::memcpy(&v, b.data() + off, sizeof(T));
    return v;
  }
  bool sig(uint64_t off, const char* s) const {
    return off + 4 <= b.size() && std::memcmp(b.data() + off, s, 4) == 0;
  }

  explicit Reader(const std::string& fname) {
    std::ifstream f(fname, std::ios::binary);
    if (!f) GAUXC_GENERIC_EXCEPTION("HDF5: cannot open " + fname);
    b.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    static const uint8_t magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (b.size() < 96 || std::memcmp(b.data(), magic, 8) != 0) GAUXC_GENERIC_EXCEPTION("HDF5: not an HDF5 file: " + fname);
    if (b[8] != 0) GAUXC_GENERIC_EXCEPTION("HDF5: only superblock version 0 is supported");
    if (b[13] != 8 || b[14] != 8) GAUXC_GENERIC_EXCEPTION("HDF5: only 8-byte offsets / lengths are supported");
    root_ohdr = get<uint64_t>(56 + 8);  // root symbol table entry at byte 56: name offset, object header
  }

  struct Msg {
    uint16_t type;
    uint64_t body;
    uint16_t size;
  };
  std::vector<Msg> messages(uint64_t addr) const {
    if (get<uint8_t>(addr) != 1) GAUXC_GENERIC_EXCEPTION("HDF5: only version 1 object headers are supported");
    const uint16_t nmsgs = get<uint16_t>(addr + 2);
    const uint32_t hsize = get<uint32_t>(addr + 8);
    std::vector<Msg> out;
    std::vector<std::pair<uint64_t, uint64_t>> blocks{{addr + 16, hsize}};
    for (size_t ib = 0; ib < blocks.size() && out.size() < nmsgs; ++ib) {
      uint64_t p = blocks[ib].first;
      const uint64_t end = p + blocks[ib].second;
      while (p + 8 <= end && out.size() < nmsgs) {
        const uint16_t t = get<uint16_t>(p), sz = get<uint16_t>(p + 2);
        if (t == 0x10) blocks.push_back({get<uint64_t>(p + 8), get<uint64_t>(p + 16)});
        out.push_back({t, p + 8, sz});
        p += 8 + sz;
      }
    }
    return out;
  }

  // name -> object header address of the members of a group
  std::map<std::string, uint64_t> children(uint64_t ohdr) const {
    std::map<std::string, uint64_t> out;
    for (auto& m : messages(ohdr)) {
      if (m.type != 0x11) continue;
      const uint64_t btree = get<uint64_t>(m.body), heap = get<uint64_t>(m.body + 8);
      if (!sig(heap, "HEAP")) GAUXC_GENERIC_EXCEPTION("HDF5: bad local heap");
      const uint64_t data = get<uint64_t>(heap + 24);
      std::vector<uint64_t> nodes{btree};
      while (!nodes.empty()) {
        const uint64_t node = nodes.back();
        nodes.pop_back();
        if (!sig(node, "TREE")) GAUXC_GENERIC_EXCEPTION("HDF5: bad B-tree node");
        const uint8_t level = get<uint8_t>(node + 5);
        const uint16_t nent = get<uint16_t>(node + 6);
        for (uint16_t i = 0; i < nent; ++i) {
          const uint64_t child = get<uint64_t>(node + 24 + 8 + 16 * (uint64_t)i);
          if (level > 0) {
            nodes.push_back(child);
            continue;
          }
          if (!sig(child, "SNOD")) GAUXC_GENERIC_EXCEPTION("HDF5: bad symbol table node");
          const uint16_t nsym = get<uint16_t>(child + 6);
          for (uint16_t k = 0; k < nsym; ++k) {
            const uint64_t e = child + 8 + 40 * (uint64_t)k;
            const uint64_t noff = get<uint64_t>(e), oh = get<uint64_t>(e + 8);
            const char* s = (const char*)b.data() + data + noff;
            out[std::string(s, strnlen(s, b.size() - (data + noff)))] = oh;
          }
        }
      }
      return out;
    }
    GAUXC_GENERIC_EXCEPTION("HDF5: not a group");
  }

  uint64_t resolve(const std::string& path) const {
    uint64_t oh = root_ohdr;
    size_t p = 0;
    while (p < path.size()) {
      while (p < path.size() && path[p] == '/') ++p;
      size_t q = path.find('/', p);
      if (q == std::string::npos) q = path.size();
      if (q == p) break;
      const std::string part = path.substr(p, q - p);
      auto ch = children(oh);
      auto it = ch.find(part);
      if (it == ch.end()) GAUXC_GENERIC_EXCEPTION("HDF5: no such object: " + path);
      oh = it->second;
      p = q;
    }
    return oh;
  }

  // ---- datatypes -----------------------------------------------------------------------------------
  struct Member {
    std::string name;
    uint32_t offset = 0;
    uint8_t cls = 0;      // class of the (array base) element
    uint32_t esize = 0;   // element size
    uint64_t count = 1;   // array elements
  };
  struct Dtype {
    uint8_t cls = 0;
    uint32_t size = 0;
    uint64_t count = 1;  // array: number of base elements
    uint8_t base_cls = 0;
    uint32_t base_size = 0;
    std::vector<Member> members;
  };
  // parses the datatype message at `p`, returns the position just after it
  uint64_t parse_dtype(uint64_t p, Dtype& d) const {
    const uint8_t b0 = get<uint8_t>(p);
    d.cls = b0 & 0x0f;
    const uint8_t ver = b0 >> 4;
    const uint32_t bits = get<uint8_t>(p + 1) | (get<uint8_t>(p + 2) << 8) | (get<uint8_t>(p + 3) << 16);
    d.size = get<uint32_t>(p + 4);
    uint64_t q = p + 8;
    switch (d.cls) {
      case 0: return q + 4;   // fixed point: bit offset, precision
      case 1: return q + 12;  // floating point
      case 10: {              // array
        const uint8_t rank = get<uint8_t>(q);
        q += (ver >= 3) ? 1 : 4;
        d.count = 1;
        for (uint8_t r = 0; r < rank; ++r) d.count *= get<uint32_t>(q + 4 * (uint64_t)r);
        q += 4 * (uint64_t)rank;
        if (ver < 3) q += 4 * (uint64_t)rank;  // permutation indices
        Dtype base;
        q = parse_dtype(q, base);
        d.base_cls = base.cls;
        d.base_size = base.size;
        return q;
      }
      case 6: {  // compound
        const uint32_t nmem = bits & 0xffff;
        for (uint32_t m = 0; m < nmem; ++m) {
          Member mem;
          const char* s = (const char*)b.data() + q;
          const size_t len = strnlen(s, b.size() - q);
          mem.name.assign(s, len);
          if (ver >= 3) {
            q += len + 1;
            const int nb = d.size < 256 ? 1 : (d.size < 65536 ? 2 : 4);
            mem.offset = 0;
            for (int k = 0; k < nb; ++k) mem.offset |= (uint32_t)get<uint8_t>(q + k) << (8 * k);
            q += nb;
          } else {
            q += (len + 1 + 7) / 8 * 8;
            mem.offset = get<uint32_t>(q);
            q += 4;
            if (ver == 1) q += 1 + 3 + 4 + 4 + 16;  // dimensionality, reserved, permutation, reserved, 4 dim sizes
          }
          Dtype mt;
          q = parse_dtype(q, mt);
          if (mt.cls == 10) { mem.cls = mt.base_cls; mem.esize = mt.base_size; mem.count = mt.count; }
          else { mem.cls = mt.cls; mem.esize = mt.size; mem.count = 1; }
          d.members.push_back(mem);
        }
        return q;
      }
      default: GAUXC_GENERIC_EXCEPTION("HDF5: unsupported datatype class " + std::to_string(d.cls));
    }
  }

  struct Dataset {
    std::vector<size_t> dims;
    Dtype type;
    uint64_t addr = UNDEF, size = 0;
  };
  Dataset dataset(const std::string& path) const {
    Dataset ds;
    for (auto& m : messages(resolve(path))) {
      if (m.type == 0x01) {
        const uint8_t ver = get<uint8_t>(m.body), rank = get<uint8_t>(m.body + 1);
        const uint64_t off = m.body + (ver == 1 ? 8 : 4);
        for (uint8_t r = 0; r < rank; ++r) ds.dims.push_back((size_t)get<uint64_t>(off + 8 * (uint64_t)r));
      } else if (m.type == 0x03) {
        parse_dtype(m.body, ds.type);
      } else if (m.type == 0x08) {
        const uint8_t ver = get<uint8_t>(m.body), cls = get<uint8_t>(m.body + 1);
        if (ver != 3) GAUXC_GENERIC_EXCEPTION("HDF5: only data layout version 3 is supported");
        if (cls == 1) { ds.addr = get<uint64_t>(m.body + 2); ds.size = get<uint64_t>(m.body + 10); }
        else if (cls == 0) { ds.size = get<uint16_t>(m.body + 2); ds.addr = m.body + 4; }
        else GAUXC_GENERIC_EXCEPTION("HDF5: chunked datasets are not supported");
      }
    }
    if (ds.type.size == 0) GAUXC_GENERIC_EXCEPTION("HDF5: not a dataset: " + path);
    return ds;
  }
  size_t nelem(const Dataset& ds) const {
    size_t n = 1;
    for (auto d : ds.dims) n *= d;
    return n;
  }
  double number(uint64_t p, uint8_t cls, uint32_t esize) const {
    if (cls == 1) return esize == 8 ? get<double>(p) : (double)get<float>(p);
    switch (esize) {
      case 8: return (double)get<int64_t>(p);
      case 4: return (double)get<int32_t>(p);
      case 2: return (double)get<int16_t>(p);
      default: return (double)get<int8_t>(p);
    }
  }
};

const Reader::Member& member(const Reader::Dtype& t, const char* name) {
  for (auto& m : t.members)
    if (m.name == name) return m;
  GAUXC_GENERIC_EXCEPTION(std::string("HDF5: compound member missing: ") + name);
}

// ---------------------------------------------------------------------------------------------------
// writer: appends one dataset to a flat root group; the whole file is rebuilt on every call (records are
// small), existing datasets are carried over byte for byte
// ---------------------------------------------------------------------------------------------------
struct Blob {
  std::vector<uint8_t> d;
  template <typename T>
  void put(T v) {
    const uint8_t* p = (const uint8_t*)&v;
    d.insert(d.end(), p, p + sizeof(T));
  }
  void bytes(const void* p, size_t n) { d.insert(d.end(), (const uint8_t*)p, (const uint8_t*)p + n); }
  void zeros(size_t n) { d.insert(d.end(), n, 0); }
  void align(size_t a) { zeros((a - d.size() % a) % a); }
  template <typename T>
  void set(size_t off, T v) { std::memcpy(d.data() + off, &v, sizeof(T)); }
};

struct OutDataset {
  std::string name;
  std::vector<uint8_t> dataspace, datatype, data;  // message bodies + raw data
};

std::vector<uint8_t> dataspace_msg(const std::vector<size_t>& dims) {
  Blob m;
  m.put<uint8_t>(1);                       // version
  m.put<uint8_t>((uint8_t)dims.size());    // rank
  m.put<uint8_t>(1);                       // flags: max dims present
  m.zeros(5);
  for (auto d : dims) m.put<uint64_t>(d);
  for (auto d : dims) m.put<uint64_t>(d);
  return m.d;
}
void put_double_type(Blob& m) {  // IEEE little-endian FP64 (the fixture's 11 20 3f 00 08 00 00 00 00 00 40 00 34 0b 00 34 ff 03 00 00)
  const uint8_t t[20] = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0, 0, 0, 0x40, 0, 0x34, 0x0b, 0, 0x34, 0xff, 0x03, 0, 0};
  m.bytes(t, 20);
}
void put_int32_type(Blob& m) {  // signed little-endian 32-bit (10 08 00 00 04 00 00 00 00 00 20 00)
  const uint8_t t[12] = {0x10, 0x08, 0, 0, 4, 0, 0, 0, 0, 0, 0x20, 0};
  m.bytes(t, 12);
}
void put_name8(Blob& m, const char* s) {
  const size_t len = std::strlen(s) + 1;
  m.bytes(s, len);
  m.zeros((8 - len % 8) % 8);
}
void put_array_member(Blob& m, const char* name, uint32_t off, uint32_t count) {  // compound v2 member, array v2 of FP64
  put_name8(m, name);
  m.put<uint32_t>(off);
  m.put<uint8_t>(0x2a); m.put<uint8_t>(0); m.put<uint8_t>(0); m.put<uint8_t>(0);  // class 10, version 2
  m.put<uint32_t>(8 * count);
  m.put<uint8_t>(1); m.zeros(3);  // rank
  m.put<uint32_t>(count);
  m.put<uint32_t>(0);             // permutation
  put_double_type(m);
}

std::vector<uint8_t> build_file(const std::vector<OutDataset>& dsets_in) {
  std::vector<OutDataset> dsets = dsets_in;
  std::sort(dsets.begin(), dsets.end(), [](const OutDataset& a, const OutDataset& b) { return a.name < b.name; });
  Blob f;
  // ---- superblock v0 (root symbol table entry filled in below) ----
  const uint8_t magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  f.bytes(magic, 8);
  const uint8_t vers[8] = {0, 0, 0, 0, 0, 8, 8, 0};
  f.bytes(vers, 8);
  f.put<uint16_t>(4);   // group leaf node K
  f.put<uint16_t>(16);  // group internal node K
  f.put<uint32_t>(0);   // consistency flags
  f.put<uint64_t>(0);       // base address
  f.put<uint64_t>(UNDEF);   // free-space info
  const size_t eof_pos = f.d.size();
  f.put<uint64_t>(0);       // end of file (patched)
  f.put<uint64_t>(UNDEF);   // driver info
  const size_t root_ste = f.d.size();
  f.zeros(40);
  // ---- root group object header ----
  const uint64_t root_oh = f.d.size();
  f.put<uint8_t>(1); f.put<uint8_t>(0); f.put<uint16_t>(1); f.put<uint32_t>(1); f.put<uint32_t>(24); f.zeros(4);
  f.put<uint16_t>(0x11); f.put<uint16_t>(16); f.put<uint8_t>(0); f.zeros(3);
  const size_t stab_pos = f.d.size();
  f.put<uint64_t>(0); f.put<uint64_t>(0);
  // ---- local heap: names ----
  Blob names;
  names.zeros(8);  // the empty name at offset 0
  std::vector<uint64_t> name_off;
  for (auto& d : dsets) {
    name_off.push_back(names.d.size());
    put_name8(names, d.name.c_str());
  }
  const uint64_t heap_free = names.d.size();
  names.put<uint64_t>(1);   // free block: next = 1 (none)
  names.put<uint64_t>(32);  // free block size
  names.zeros(16);
  f.align(8);
  const uint64_t heap = f.d.size();
  f.bytes("HEAP", 4); f.put<uint8_t>(0); f.zeros(3);
  f.put<uint64_t>(names.d.size());
  f.put<uint64_t>(heap_free);
  f.put<uint64_t>(heap + 32);
  f.bytes(names.d.data(), names.d.size());
  // ---- dataset object headers + data ----
  std::vector<uint64_t> oh_addr;
  for (auto& d : dsets) {
    f.align(8);
    // raw data first
    const uint64_t data_addr = d.data.empty() ? UNDEF : f.d.size();
    f.bytes(d.data.data(), d.data.size());
    f.align(8);
    oh_addr.push_back(f.d.size());
    Blob msgs;
    auto add = [&](uint16_t type, const std::vector<uint8_t>& body, uint8_t flags) {
      const size_t padded = (body.size() + 7) / 8 * 8;
      msgs.put<uint16_t>(type); msgs.put<uint16_t>((uint16_t)padded); msgs.put<uint8_t>(flags); msgs.zeros(3);
      msgs.bytes(body.data(), body.size());
      msgs.zeros(padded - body.size());
    };
    add(0x01, d.dataspace, 0);
    add(0x03, d.datatype, 1);                                   // constant
    add(0x05, std::vector<uint8_t>{2, 2, 2, 1, 0, 0, 0, 0}, 1);  // fill value v2: late allocation, as the fixtures
    Blob lay;
    lay.put<uint8_t>(3); lay.put<uint8_t>(1); lay.put<uint64_t>(data_addr); lay.put<uint64_t>(d.data.size());
    add(0x08, lay.d, 0);
    f.put<uint8_t>(1); f.put<uint8_t>(0); f.put<uint16_t>(4); f.put<uint32_t>(1); f.put<uint32_t>((uint32_t)msgs.d.size());
    f.zeros(4);
    f.bytes(msgs.d.data(), msgs.d.size());
  }
  // ---- symbol table nodes (<= 8 entries each) + B-tree ----
  std::vector<uint64_t> snods, last_name;
  for (size_t i = 0; i < dsets.size(); i += 8) {
    f.align(8);
    snods.push_back(f.d.size());
    const size_t n = std::min<size_t>(8, dsets.size() - i);
    f.bytes("SNOD", 4); f.put<uint8_t>(1); f.put<uint8_t>(0); f.put<uint16_t>((uint16_t)n);
    for (size_t k = 0; k < 8; ++k) {
      if (k < n) {
        f.put<uint64_t>(name_off[i + k]); f.put<uint64_t>(oh_addr[i + k]); f.put<uint32_t>(0); f.put<uint32_t>(0);
        f.zeros(16);
      } else {
        f.zeros(40);
      }
    }
    last_name.push_back(name_off[i + n - 1]);
  }
  if (snods.size() > 32) GAUXC_GENERIC_EXCEPTION("HDF5: too many datasets for one B-tree node");
  f.align(8);
  const uint64_t btree = f.d.size();
  f.bytes("TREE", 4); f.put<uint8_t>(0); f.put<uint8_t>(0); f.put<uint16_t>((uint16_t)snods.size());
  f.put<uint64_t>(UNDEF); f.put<uint64_t>(UNDEF);
  f.put<uint64_t>(0);  // key 0: the empty name
  for (size_t i = 0; i < snods.size(); ++i) { f.put<uint64_t>(snods[i]); f.put<uint64_t>(last_name[i]); }
  f.zeros((2 * 16 - snods.size()) * 16);  // room of a full node (2K children)
  // ---- patch ----
  f.set<uint64_t>(stab_pos, btree);
  f.set<uint64_t>(stab_pos + 8, heap);
  f.set<uint64_t>(root_ste, 0);
  f.set<uint64_t>(root_ste + 8, root_oh);
  f.set<uint32_t>(root_ste + 16, 1);  // cache type 1: scratch = B-tree and heap addresses
  f.set<uint64_t>(root_ste + 24, btree);
  f.set<uint64_t>(root_ste + 32, heap);
  f.set<uint64_t>(eof_pos, f.d.size());
  return f.d;
}

// existing datasets of a flat file, re-encoded from their stored message bodies
std::vector<OutDataset> load_existing(const std::string& fname) {
  std::vector<OutDataset> out;
  std::ifstream probe(fname, std::ios::binary);
  if (!probe) return out;
  probe.close();
  Reader r(fname);
  for (auto& kv : r.children(r.root_ohdr)) {
    OutDataset d;
    d.name = kv.first;
    uint64_t addr = UNDEF, size = 0;
    bool is_dataset = false;
    for (auto& m : r.messages(kv.second)) {
      if (m.type == 0x01) d.dataspace.assign(r.b.begin() + m.body, r.b.begin() + m.body + m.size);
      if (m.type == 0x03) { d.datatype.assign(r.b.begin() + m.body, r.b.begin() + m.body + m.size); is_dataset = true; }
      if (m.type == 0x08 && r.get<uint8_t>(m.body + 1) == 1) { addr = r.get<uint64_t>(m.body + 2); size = r.get<uint64_t>(m.body + 10); }
    }
    if (!is_dataset) GAUXC_GENERIC_EXCEPTION("HDF5: cannot rewrite a file with sub-groups: " + fname);
    if (addr != UNDEF) d.data.assign(r.b.begin() + addr, r.b.begin() + addr + size);
    out.push_back(std::move(d));
  }
  return out;
}

void add_and_write(const std::string& fname, OutDataset nd) {
  auto all = load_existing(fname);
  std::string name = nd.name;
  while (!name.empty() && name[0] == '/') name.erase(0, 1);
  if (name.empty() || name.find('/') != std::string::npos) GAUXC_GENERIC_EXCEPTION("HDF5: datasets are written to the root group only");
  nd.name = name;
  for (auto& d : all)
    if (d.name == name) GAUXC_GENERIC_EXCEPTION("Dataset Creation Failed");  // H5Dcreate on an existing name
  all.push_back(std::move(nd));
  const auto bytes = build_file(all);
  std::ofstream f(fname, std::ios::binary | std::ios::trunc);
  if (!f) GAUXC_GENERIC_EXCEPTION("HDF5: cannot write " + fname);
  f.write((const char*)bytes.data(), (std::streamsize)bytes.size());
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
void read_hdf5_record(Molecule& mol, const std::string& fname, const std::string& dset) {
  // hdf5_read.cxx:103-156: compound {Atomic Number, X / Y / Z Coordinate}
  Reader r(fname);
  const auto ds = r.dataset(dset);
  if (ds.type.cls != 6) GAUXC_GENERIC_EXCEPTION("HDF5: " + dset + " is not a compound dataset");
  if (ds.dims.size() != 1) GAUXC_GENERIC_EXCEPTION("Dataspace for Molecule Record Must Be 1D");
  const auto &mz = member(ds.type, "Atomic Number"), &mx = member(ds.type, "X Coordinate"),
             &my = member(ds.type, "Y Coordinate"), &mzc = member(ds.type, "Z Coordinate");
  mol.clear();
  for (size_t i = 0; i < ds.dims[0]; ++i) {
    const uint64_t p = ds.addr + i * ds.type.size;
    mol.push_back({(int64_t)r.number(p + mz.offset, mz.cls, mz.esize), r.number(p + mx.offset, mx.cls, mx.esize),
                   r.number(p + my.offset, my.cls, my.esize), r.number(p + mzc.offset, mzc.cls, mzc.esize)});
  }
}

void read_hdf5_record(BasisSet& basis, const std::string& fname, const std::string& dset) {
  // hdf5_read.cxx:47-100: compound {NPRIM, L, PURE, ALPHA[], COEFF[], ORIGIN[3]}; coefficients are stored normalised
  Reader r(fname);
  const auto ds = r.dataset(dset);
  if (ds.type.cls != 6) GAUXC_GENERIC_EXCEPTION("HDF5: " + dset + " is not a compound dataset");
  if (ds.dims.size() != 1) GAUXC_GENERIC_EXCEPTION("Dataspace for Basis Record Must Be 1D");
  const auto &mn = member(ds.type, "NPRIM"), &ml = member(ds.type, "L"), &mp = member(ds.type, "PURE"),
             &ma = member(ds.type, "ALPHA"), &mc = member(ds.type, "COEFF"), &mo = member(ds.type, "ORIGIN");
  basis.clear();
  for (size_t i = 0; i < ds.dims[0]; ++i) {
    const uint64_t p = ds.addr + i * ds.type.size;
    const int nprim = (int)r.number(p + mn.offset, mn.cls, mn.esize);
    if (nprim < 0 || (uint64_t)nprim > ma.count || (uint64_t)nprim > mc.count || nprim > 32)
      GAUXC_GENERIC_EXCEPTION("HDF5: invalid NPRIM in " + dset);
    double alpha[32] = {0}, coeff[32] = {0}, O[3];
    for (int k = 0; k < nprim; ++k) {
      alpha[k] = r.number(p + ma.offset + 8 * (uint64_t)k, ma.cls, ma.esize);
      coeff[k] = r.number(p + mc.offset + 8 * (uint64_t)k, mc.cls, mc.esize);
    }
    for (int k = 0; k < 3; ++k) O[k] = r.number(p + mo.offset + 8 * (uint64_t)k, mo.cls, mo.esize);
    basis.push_back(Shell(nprim, (int)r.number(p + ml.offset, ml.cls, ml.esize),
                          (int)r.number(p + mp.offset, mp.cls, mp.esize), alpha, coeff, O, /*normalize=*/false));
  }
}

void write_hdf5_record(const Molecule& mol, const std::string& fname, const std::string& dset) {
  OutDataset d;
  d.name = dset;
  d.dataspace = dataspace_msg({mol.size()});
  Blob t;  // compound v1, 4 members, 32 bytes: the layout of the reference's Atom struct
  t.put<uint8_t>(0x16); t.put<uint8_t>(4); t.put<uint8_t>(0); t.put<uint8_t>(0);
  t.put<uint32_t>(32);
  auto member_v1 = [&](const char* name, uint32_t off, bool integer) {
    put_name8(t, name);
    t.put<uint32_t>(off);
    t.zeros(1 + 3 + 4 + 4 + 16);
    if (integer) put_int32_type(t);
    else put_double_type(t);
  };
  member_v1("Atomic Number", 0, true);
  member_v1("X Coordinate", 8, false);
  member_v1("Y Coordinate", 16, false);
  member_v1("Z Coordinate", 24, false);
  d.datatype = t.d;
  Blob data;
  for (auto& a : mol) {
    data.put<int64_t>(a.Z);  // H5T_NATIVE_INT reads the low 4 bytes (little endian)
    data.put<double>(a.x); data.put<double>(a.y); data.put<double>(a.z);
  }
  d.data = data.d;
  add_and_write(fname, std::move(d));
}

void write_hdf5_record(const BasisSet& basis, const std::string& fname, const std::string& dset) {
  OutDataset d;
  d.name = dset;
  d.dataspace = dataspace_msg({basis.size()});
  Blob t;  // compound v2, 6 members, 552 bytes: shell_t of hdf5_util.hpp (arrays of 32 in memory, 16 declared)
  t.put<uint8_t>(0x26); t.put<uint8_t>(6); t.put<uint8_t>(0); t.put<uint8_t>(0);
  t.put<uint32_t>(552);
  auto int_member = [&](const char* name, uint32_t off) {
    put_name8(t, name);
    t.put<uint32_t>(off);
    put_int32_type(t);
  };
  int_member("NPRIM", 0);
  int_member("L", 4);
  int_member("PURE", 8);
  put_array_member(t, "ALPHA", 16, 16);
  put_array_member(t, "COEFF", 272, 16);
  put_array_member(t, "ORIGIN", 528, 3);
  d.datatype = t.d;
  Blob data;
  for (auto& sh : basis) {
    if (sh.nprim > 16) GAUXC_GENERIC_EXCEPTION("HDF5 basis records hold at most 16 primitives per shell");
    data.put<int32_t>(sh.nprim); data.put<int32_t>(sh.l); data.put<int32_t>(sh.pure); data.put<int32_t>(0);
    for (int k = 0; k < 32; ++k) data.put<double>(sh.alpha[k]);
    for (int k = 0; k < 32; ++k) data.put<double>(sh.coeff[k]);
    for (int k = 0; k < 3; ++k) data.put<double>(sh.O[k]);
  }
  d.data = data.d;
  add_and_write(fname, std::move(d));
}

void read_hdf5_dataset(const std::string& fname, const std::string& dset, std::vector<double>& data,
                       std::vector<size_t>& dims) {
  Reader r(fname);
  const auto ds = r.dataset(dset);
  if (ds.type.cls > 1) GAUXC_GENERIC_EXCEPTION("HDF5: " + dset + " is not a numeric dataset");
  dims = ds.dims;
  const size_t n = r.nelem(ds);
  data.resize(n);
  if (n && ds.addr == UNDEF) GAUXC_GENERIC_EXCEPTION("HDF5: dataset has no storage: " + dset);
  for (size_t i = 0; i < n; ++i) data[i] = r.number(ds.addr + i * ds.type.size, ds.type.cls, ds.type.size);
}

void write_hdf5_dataset(const std::string& fname, const std::string& dset, const double* data,
                        const std::vector<size_t>& dims) {
  OutDataset d;
  d.name = dset;
  d.dataspace = dataspace_msg(dims);
  Blob t;
  put_double_type(t);
  d.datatype = t.d;
  size_t n = 1;
  for (auto x : dims) n *= x;
  d.data.assign((const uint8_t*)data, (const uint8_t*)data + n * sizeof(double));
  add_and_write(fname, std::move(d));
}

}  // namespace GauXC
