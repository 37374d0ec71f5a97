// Native HDF5 subset for particle files ("next" row 2): the on-disk layouts of
//   ClusterParticles.write_particles / from_file       (particles.py:239-274, 337-363)
//   ClusterParticles.write_to_gadget_file / from_gadget_file (particles.py:280-335, 453-552)
// h5py / libhdf5 are not available, so this restates the published HDF5 File Format
// Specification (version 1.x "earliest" layout, which is what h5py's default libver
// writes): superblock v0, symbol-table groups (v1 B-tree + SNOD + local heap), v1 object
// headers, contiguous little-endian datasets, attribute messages, variable-length UTF-8
// string attributes in a global heap collection.
//
// Writer: dataset bytes stream to the file as they arrive (from host memory, or straight
// from device planes through a pack kernel and two pinned staging slots so that the file
// write of chunk k overlaps the pack + D2H copy of chunk k+1); all metadata is kept in memory
// and laid out behind the data when the file is closed.
// Reader: superblock v0/v1 at any user-block offset, multi-level group B-trees, object header
// continuation blocks, dataspace v1/v2, layout v1-v4 (contiguous/compact), attribute v1-v3.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <errno.h>
#include <string.h>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "cgb_common.cuh"

#define CGB_H5_CHECK(expr)                 \
    do {                                   \
        const int _rc = (expr);            \
        if (_rc != CGB_OK) return _rc;     \
    } while (0)

#define CGB_H5_IOFAIL(...)                 \
    do {                                   \
        cgb_set_error(__VA_ARGS__);        \
        return CGB_E_IO;                   \
    } while (0)

namespace {

const uint64_t UNDEF = ~0ull;
const uint8_t H5_SIG[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
const int LEAF_K = 4;       // symbol-table node holds <= 2*LEAF_K entries
const int INTERNAL_K = 16;  // group B-tree node holds <= 2*INTERNAL_K children
const size_t STAGE_BYTES = 32u << 20;

size_t dtype_size(int dt)
{
    switch (dt) {
    case CGB_H5_F32: case CGB_H5_I32: case CGB_H5_U32: return 4;
    case CGB_H5_F64: case CGB_H5_I64: case CGB_H5_U64: return 8;
    default: return 0;
    }
}

struct Attr {
    std::string name;
    int dtype = 0;        // CGB_H5_*; 0 = present in the file but of a type outside the subset
    int64_t count = -1;   // -1: scalar dataspace; else 1-D length (STR: always scalar)
    std::vector<uint8_t> data;  // numeric bytes, or the string without terminator
};

struct Node {
    bool is_group = true;
    std::map<std::string, std::unique_ptr<Node>> kids;  // strcmp order = symbol-table order
    std::vector<Attr> attrs;
    // datasets
    int dtype = 0;
    int rank = 0;
    int64_t dims[4] = {0, 0, 0, 0};
    uint64_t data_addr = UNDEF;
    uint64_t data_bytes = 0;
    std::vector<uint8_t> compact;  // reader: compact-layout payload
    // writer: addresses assigned at close
    uint64_t hdr_addr = 0, btree_addr = 0, heap_addr = 0;
};

}  // namespace

struct cgb_h5_file {
    int fd = -1;
    bool writing = false;
    std::string path;
    Node root;
    uint64_t cursor = 0;  // writer: next free data byte
    // reader
    const uint8_t* map = nullptr;
    size_t map_len = 0;
    uint64_t base = 0;
    // device staging (allocated on first device write)
    void* pinned[2] = {nullptr, nullptr};
    void* d_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

namespace {

// ---------------------------------------------------------------------------
// tree navigation
// ---------------------------------------------------------------------------
std::vector<std::string> split_path(const char* path)
{
    std::vector<std::string> parts;
    std::string cur;
    for (const char* p = path; *p; ++p) {
        if (*p == '/') {
            if (!cur.empty()) parts.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(*p);
        }
    }
    if (!cur.empty()) parts.push_back(cur);
    return parts;
}

Node* find_node(cgb_h5_file* f, const char* path)
{
    Node* n = &f->root;
    for (const std::string& part : split_path(path)) {
        if (!n->is_group) return nullptr;
        auto it = n->kids.find(part);
        if (it == n->kids.end()) return nullptr;
        n = it->second.get();
    }
    return n;
}

// parent group + leaf name for a path that is about to be created
int find_parent(cgb_h5_file* f, const char* path, Node** parent, std::string* leaf)
{
    std::vector<std::string> parts = split_path(path);
    CGB_REQUIRE(!parts.empty(), "cgb_h5: empty object path");
    Node* n = &f->root;
    for (size_t i = 0; i + 1 < parts.size(); ++i) {
        auto it = n->kids.find(parts[i]);
        CGB_REQUIRE(it != n->kids.end() && it->second->is_group, "cgb_h5: group '%s' of '%s' does not exist",
                    parts[i].c_str(), path);
        n = it->second.get();
    }
    CGB_REQUIRE(n->kids.find(parts.back()) == n->kids.end(), "cgb_h5: '%s' already exists", path);
    *parent = n;
    *leaf = parts.back();
    return CGB_OK;
}

int pwrite_all(int fd, const void* buf, size_t len, uint64_t off)
{
    const uint8_t* p = (const uint8_t*)buf;
    while (len > 0) {
        const ssize_t w = pwrite(fd, p, len, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            CGB_H5_IOFAIL("cgb_h5: write failed: %s", strerror(errno));
        }
        p += w;
        off += (uint64_t)w;
        len -= (size_t)w;
    }
    return CGB_OK;
}

uint64_t align_up64(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// little-endian byte buffer addressed by absolute file offset
// ---------------------------------------------------------------------------
struct Meta {
    uint64_t start;
    std::vector<uint8_t> b;
    explicit Meta(uint64_t s) : start(s) {}
    uint64_t alloc(size_t n)  // 8-byte aligned, zero-filled
    {
        const size_t at = (b.size() + 7) / 8 * 8;
        b.resize(at + n, 0);
        return start + at;
    }
    uint8_t* at(uint64_t addr) { return b.data() + (addr - start); }
};

void put(std::vector<uint8_t>& v, uint64_t x, int bytes)
{
    for (int i = 0; i < bytes; ++i) v.push_back((uint8_t)(x >> (8 * i)));
}
void put_at(uint8_t* p, uint64_t x, int bytes)
{
    for (int i = 0; i < bytes; ++i) p[i] = (uint8_t)(x >> (8 * i));
}
void pad8(std::vector<uint8_t>& v)
{
    while (v.size() % 8) v.push_back(0);
}

// datatype message body (spec IV.A.2.d)
std::vector<uint8_t> encode_dtype(int dt)
{
    std::vector<uint8_t> v;
    switch (dt) {
    case CGB_H5_F32:
        v = {0x11, 0x20, 0x1f, 0x00, 4, 0, 0, 0, 0, 0, 32, 0, 23, 8, 0, 23, 127, 0, 0, 0};
        break;
    case CGB_H5_F64:
        v = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xff, 0x03, 0, 0};
        break;
    case CGB_H5_I32: v = {0x10, 0x08, 0, 0, 4, 0, 0, 0, 0, 0, 32, 0}; break;
    case CGB_H5_U32: v = {0x10, 0x00, 0, 0, 4, 0, 0, 0, 0, 0, 32, 0}; break;
    case CGB_H5_I64: v = {0x10, 0x08, 0, 0, 8, 0, 0, 0, 0, 0, 64, 0}; break;
    case CGB_H5_U64: v = {0x10, 0x00, 0, 0, 8, 0, 0, 0, 0, 0, 64, 0}; break;
    case CGB_H5_STR:
        // variable-length string, null-terminated, UTF-8; base type = 1-byte unsigned integer
        v = {0x19, 0x01, 0x01, 0x00, 16, 0, 0, 0, 0x10, 0x00, 0, 0, 1, 0, 0, 0, 0, 0, 8, 0};
        break;
    }
    return v;
}

// dataspace message v1 (spec IV.A.2.b)
std::vector<uint8_t> encode_space(int rank, const int64_t* dims)
{
    std::vector<uint8_t> v = {1, (uint8_t)rank, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < rank; ++i) put(v, (uint64_t)dims[i], 8);
    return v;
}

struct VlenRef {
    uint64_t gcol_addr;
    uint32_t index;
};

// attribute message v1 (spec IV.A.2.m)
std::vector<uint8_t> encode_attr(const Attr& a, const VlenRef* vl)
{
    std::vector<uint8_t> dt = encode_dtype(a.dtype);
    const int64_t dim = a.count;
    std::vector<uint8_t> sp = encode_space(a.count < 0 ? 0 : 1, &dim);
    std::vector<uint8_t> v = {1, 0};
    put(v, a.name.size() + 1, 2);
    put(v, dt.size(), 2);
    put(v, sp.size(), 2);
    v.insert(v.end(), a.name.begin(), a.name.end());
    v.push_back(0);
    pad8(v);
    v.insert(v.end(), dt.begin(), dt.end());
    pad8(v);
    v.insert(v.end(), sp.begin(), sp.end());
    pad8(v);
    if (a.dtype == CGB_H5_STR) {
        put(v, a.data.size(), 4);
        put(v, a.data.empty() ? 0 : vl->gcol_addr, 8);
        put(v, a.data.empty() ? 0 : vl->index, 4);
    } else {
        v.insert(v.end(), a.data.begin(), a.data.end());
    }
    return v;
}

struct Msg {
    uint16_t type;
    uint8_t flags;
    std::vector<uint8_t> body;
};

// v1 object header with a single chunk (spec IV.A.1.a)
uint64_t write_object_header(Meta& m, std::vector<Msg>& msgs)
{
    size_t total = 0;
    for (Msg& g : msgs) {
        pad8(g.body);
        total += 8 + g.body.size();
    }
    const uint64_t addr = m.alloc(16 + total);
    uint8_t* p = m.at(addr);
    p[0] = 1;
    put_at(p + 2, msgs.size(), 2);
    put_at(p + 4, 1, 4);      // reference count
    put_at(p + 8, total, 4);  // bytes of message data
    p += 16;
    for (const Msg& g : msgs) {
        put_at(p, g.type, 2);
        put_at(p + 2, g.body.size(), 2);
        p[4] = g.flags;
        memcpy(p + 8, g.body.data(), g.body.size());
        p += 8 + g.body.size();
    }
    return addr;
}

struct StringHeap {
    std::map<const Attr*, VlenRef> refs;
};

void collect_strings(const Node& n, std::vector<const Attr*>& out)
{
    for (const Attr& a : n.attrs)
        if (a.dtype == CGB_H5_STR && !a.data.empty()) out.push_back(&a);
    for (const auto& kv : n.kids) collect_strings(*kv.second, out);
}

// global heap collection (spec III.E)
int write_global_heap(Meta& m, const Node& root, StringHeap& sh)
{
    std::vector<const Attr*> strs;
    collect_strings(root, strs);
    if (strs.empty()) return CGB_OK;
    CGB_REQUIRE(strs.size() < 65535, "cgb_h5: too many string attributes (%zu)", strs.size());
    size_t used = 16;
    for (const Attr* a : strs) used += 16 + (a->data.size() + 7) / 8 * 8;
    size_t size = used < 4096 ? 4096 : used;
    if (size - used > 0 && size - used < 16) size += 8;
    const uint64_t addr = m.alloc(size);
    uint8_t* p = m.at(addr);
    memcpy(p, "GCOL", 4);
    p[4] = 1;
    put_at(p + 8, size, 8);
    uint8_t* q = p + 16;
    uint32_t idx = 1;
    for (const Attr* a : strs) {
        put_at(q, idx, 2);
        put_at(q + 2, 1, 2);  // reference count
        put_at(q + 8, a->data.size(), 8);
        memcpy(q + 16, a->data.data(), a->data.size());
        sh.refs[a] = VlenRef{addr, idx};
        q += 16 + (a->data.size() + 7) / 8 * 8;
        ++idx;
    }
    if ((size_t)(q - p) < size) put_at(q + 8, size - (size_t)(q - p), 8);  // free-space object 0
    return CGB_OK;
}

void attr_messages(const Node& n, const StringHeap& sh, std::vector<Msg>& msgs)
{
    for (const Attr& a : n.attrs) {
        const VlenRef* vl = nullptr;
        auto it = sh.refs.find(&a);
        if (it != sh.refs.end()) vl = &it->second;
        msgs.push_back(Msg{0x000C, 0, encode_attr(a, vl)});
    }
}

int serialize_node(Meta& m, Node& n, const StringHeap& sh)
{
    if (!n.is_group) {
        std::vector<Msg> msgs;
        msgs.push_back(Msg{0x0001, 0, encode_space(n.rank, n.dims)});
        msgs.push_back(Msg{0x0003, 1, encode_dtype(n.dtype)});
        msgs.push_back(Msg{0x0005, 1, {1, 2, 2, 1, 0, 0, 0, 0}});  // fill value v1: late alloc, default
        std::vector<uint8_t> lay = {3, 1};                         // layout v3, contiguous
        put(lay, n.data_bytes ? n.data_addr : UNDEF, 8);
        put(lay, n.data_bytes, 8);
        msgs.push_back(Msg{0x0008, 0, lay});
        attr_messages(n, sh, msgs);
        for (const Msg& g : msgs)
            CGB_REQUIRE(g.body.size() < 65528, "cgb_h5: attribute too large for an object header message");
        n.hdr_addr = write_object_header(m, msgs);
        return CGB_OK;
    }
    for (auto& kv : n.kids) CGB_H5_CHECK(serialize_node(m, *kv.second, sh));
    const size_t nk = n.kids.size();
    const size_t nsnod = (nk + 2 * LEAF_K - 1) / (2 * LEAF_K);
    CGB_REQUIRE(nsnod <= (size_t)(2 * INTERNAL_K), "cgb_h5: more than %d objects in one group", 4 * LEAF_K * INTERNAL_K);

    // local heap: offset 0 = "", then the link names (spec III.D)
    std::vector<uint64_t> name_off;
    std::vector<uint8_t> seg(8, 0);
    for (const auto& kv : n.kids) {
        name_off.push_back(seg.size());
        seg.insert(seg.end(), kv.first.begin(), kv.first.end());
        seg.push_back(0);
        pad8(seg);
    }
    n.heap_addr = m.alloc(32 + seg.size());
    {
        uint8_t* p = m.at(n.heap_addr);
        memcpy(p, "HEAP", 4);
        put_at(p + 8, seg.size(), 8);
        put_at(p + 16, 1, 8);  // no free block (H5HL_FREE_NULL)
        put_at(p + 24, n.heap_addr + 32, 8);
        memcpy(p + 32, seg.data(), seg.size());
    }
    // symbol-table nodes (spec III.C) in name order
    std::vector<uint64_t> snod_addr, snod_lastname;
    {
        auto it = n.kids.begin();
        size_t k = 0;
        for (size_t s = 0; s < nsnod; ++s) {
            const uint64_t a = m.alloc(8 + 2 * LEAF_K * 40);
            const size_t cnt = (nk - k) < (size_t)(2 * LEAF_K) ? nk - k : (size_t)(2 * LEAF_K);
            uint8_t* p = m.at(a);
            memcpy(p, "SNOD", 4);
            p[4] = 1;
            put_at(p + 6, cnt, 2);
            for (size_t e = 0; e < cnt; ++e, ++it, ++k) {
                uint8_t* q = m.at(a) + 8 + e * 40;
                const Node& c = *it->second;
                put_at(q, name_off[k], 8);
                put_at(q + 8, c.hdr_addr, 8);
                if (c.is_group) {
                    put_at(q + 16, 1, 4);
                    put_at(q + 24, c.btree_addr, 8);
                    put_at(q + 32, c.heap_addr, 8);
                }
            }
            snod_addr.push_back(a);
            snod_lastname.push_back(name_off[k - 1]);
        }
    }
    // group B-tree, one leaf-level node (spec III.A.1)
    n.btree_addr = m.alloc(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8);
    {
        uint8_t* p = m.at(n.btree_addr);
        memcpy(p, "TREE", 4);
        put_at(p + 6, nsnod, 2);
        put_at(p + 8, UNDEF, 8);
        put_at(p + 16, UNDEF, 8);
        uint8_t* q = p + 24;  // key0 = "" at heap offset 0
        for (size_t s = 0; s < nsnod; ++s) {
            put_at(q + 8, snod_addr[s], 8);
            put_at(q + 16, snod_lastname[s], 8);
            q += 16;
        }
    }
    std::vector<Msg> msgs;
    std::vector<uint8_t> stab;
    put(stab, n.btree_addr, 8);
    put(stab, n.heap_addr, 8);
    msgs.push_back(Msg{0x0011, 0, stab});
    attr_messages(n, sh, msgs);
    n.hdr_addr = write_object_header(m, msgs);
    return CGB_OK;
}

int finish_file(cgb_h5_file* f)
{
    Meta m(align_up64(f->cursor, 8));
    StringHeap sh;
    CGB_H5_CHECK(write_global_heap(m, f->root, sh));
    CGB_H5_CHECK(serialize_node(m, f->root, sh));
    const uint64_t eof = m.start + m.b.size();
    CGB_H5_CHECK(pwrite_all(f->fd, m.b.data(), m.b.size(), m.start));
    std::vector<uint8_t> sb(H5_SIG, H5_SIG + 8);
    const uint8_t head[8] = {0, 0, 0, 0, 0, 8, 8, 0};
    sb.insert(sb.end(), head, head + 8);
    put(sb, LEAF_K, 2);
    put(sb, INTERNAL_K, 2);
    put(sb, 0, 4);      // consistency flags
    put(sb, 0, 8);      // base address
    put(sb, UNDEF, 8);  // free-space info
    put(sb, eof, 8);
    put(sb, UNDEF, 8);  // driver info
    put(sb, 0, 8);      // root entry: link name offset
    put(sb, f->root.hdr_addr, 8);
    put(sb, 1, 4);  // cached symbol-table addresses follow
    put(sb, 0, 4);
    put(sb, f->root.btree_addr, 8);
    put(sb, f->root.heap_addr, 8);
    CGB_H5_CHECK(pwrite_all(f->fd, sb.data(), sb.size(), 0));
    return CGB_OK;
}

void release(cgb_h5_file* f)
{
    for (int s = 0; s < 2; ++s) {
        if (f->pinned[s]) cudaFreeHost(f->pinned[s]);
        if (f->d_stage[s]) cudaFree(f->d_stage[s]);
        if (f->ev[s]) cudaEventDestroy(f->ev[s]);
    }
    if (f->map) munmap((void*)f->map, f->map_len);
    if (f->fd >= 0) close(f->fd);
    delete f;
}

int new_dataset(cgb_h5_file* f, const char* path, int dtype, int rank, const int64_t* dims, Node** out)
{
    CGB_REQUIRE(f && f->writing, "cgb_h5: file is not open for writing");
    CGB_REQUIRE(dtype_size(dtype) > 0, "cgb_h5: dataset dtype %d not supported", dtype);
    CGB_REQUIRE(rank >= 0 && rank <= 4, "cgb_h5: dataset rank must be 0..4 (0 = scalar)");
    Node* parent;
    std::string leaf;
    CGB_H5_CHECK(find_parent(f, path, &parent, &leaf));
    std::unique_ptr<Node> n(new Node);
    n->is_group = false;
    n->dtype = dtype;
    n->rank = rank;
    uint64_t count = 1;
    for (int i = 0; i < rank; ++i) {
        CGB_REQUIRE(dims[i] >= 0, "cgb_h5: negative dimension");
        n->dims[i] = dims[i];
        count *= (uint64_t)dims[i];
    }
    n->data_bytes = count * dtype_size(dtype);
    f->cursor = align_up64(f->cursor, n->data_bytes >= (1u << 20) ? 4096 : 8);
    n->data_addr = f->cursor;
    f->cursor += n->data_bytes;
    *out = n.get();
    parent->kids[leaf] = std::move(n);
    return CGB_OK;
}

// (ncomp, ld) planes of SrcT -> interleaved (m, ncomp) rows of DstT, scaled
template <typename SrcT, typename DstT>
__global__ void cgb_h5_pack_kernel(const SrcT* __restrict__ src, long long ld, long long first, long long m,
                                   int ncomp, double scale, DstT* __restrict__ dst)
{
    const long long total = m * ncomp;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (long long)gridDim.x * blockDim.x) {
        const long long i = j / ncomp;
        const int k = (int)(j - i * ncomp);
        dst[j] = (DstT)((double)src[k * ld + first + i] * scale);
    }
}

template <typename SrcT, typename DstT>
void launch_pack(const void* src, long long ld, long long first, long long m, int ncomp, double scale, void* dst,
                 cudaStream_t st)
{
    const long long total = m * ncomp;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cgb_h5_pack_kernel<SrcT, DstT><<<(unsigned)blocks, 256, 0, st>>>((const SrcT*)src, ld, first, m, ncomp, scale,
                                                                      (DstT*)dst);
}

// interleaved (m, ncomp) rows of SrcT -> (ncomp, ld) planes of DstT, scaled (the read direction)
template <typename SrcT, typename DstT>
__global__ void cgb_h5_unpack_kernel(const SrcT* __restrict__ src, long long ld, long long first, long long m,
                                     int ncomp, double scale, DstT* __restrict__ dst)
{
    const long long total = m * ncomp;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (long long)gridDim.x * blockDim.x) {
        const long long i = j / ncomp;
        const int k = (int)(j - i * ncomp);
        dst[k * ld + first + i] = (DstT)((double)src[j] * scale);
    }
}

template <typename SrcT, typename DstT>
void launch_unpack(const void* src, long long ld, long long first, long long m, int ncomp, double scale, void* dst,
                   cudaStream_t st)
{
    const long long total = m * ncomp;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cgb_h5_unpack_kernel<SrcT, DstT><<<(unsigned)blocks, 256, 0, st>>>((const SrcT*)src, ld, first, m, ncomp, scale,
                                                                        (DstT*)dst);
}

int ensure_staging(cgb_h5_file* f)
{
    for (int s = 0; s < 2; ++s) {
        if (!f->pinned[s]) CGB_CUDA_CHECK(cudaMallocHost(&f->pinned[s], STAGE_BYTES));
        if (!f->d_stage[s]) CGB_CUDA_CHECK(cudaMalloc(&f->d_stage[s], STAGE_BYTES));
        if (!f->ev[s]) CGB_CUDA_CHECK(cudaEventCreateWithFlags(&f->ev[s], cudaEventDisableTiming));
    }
    return CGB_OK;
}

// ---------------------------------------------------------------------------
// reader
// ---------------------------------------------------------------------------
struct Reader {
    const uint8_t* p;
    size_t len;
    uint64_t base;
    bool ok(uint64_t addr, uint64_t n) const { return addr != UNDEF && addr + base + n <= len && addr + base + n >= addr; }
    const uint8_t* at(uint64_t addr) const { return p + base + addr; }
};

uint64_t get(const uint8_t* p, int bytes)
{
    uint64_t x = 0;
    for (int i = 0; i < bytes; ++i) x |= (uint64_t)p[i] << (8 * i);
    return x;
}

// datatype message -> CGB_H5_* (0 if outside the subset); *fixed_len = size of a fixed string
int decode_dtype(const uint8_t* d, size_t n, size_t* fixed_len)
{
    if (n < 8) return 0;
    const int cls = d[0] & 0x0f;
    const uint32_t size = (uint32_t)get(d + 4, 4);
    const bool big_endian = d[1] & 1;
    if (cls == 0 && !big_endian) {
        const bool sgn = d[1] & 0x08;
        if (size == 4) return sgn ? CGB_H5_I32 : CGB_H5_U32;
        if (size == 8) return sgn ? CGB_H5_I64 : CGB_H5_U64;
        return 0;
    }
    if (cls == 1 && !big_endian) {
        if (size == 4) return CGB_H5_F32;
        if (size == 8) return CGB_H5_F64;
        return 0;
    }
    if (cls == 3) {
        *fixed_len = size;
        return CGB_H5_STR;
    }
    if (cls == 9 && (d[1] & 0x0f) == 1) {
        *fixed_len = 0;
        return CGB_H5_STR;
    }
    return 0;
}

// dataspace message -> rank/dims; returns false if unsupported
bool decode_space(const uint8_t* d, size_t n, int* rank, int64_t* dims, int maxrank)
{
    if (n < 4) return false;
    const int ver = d[0];
    const int r = d[1];
    size_t off;
    if (ver == 1) off = 8;
    else if (ver == 2) {
        off = 4;
        if (d[3] == 2) return false;  // null dataspace
    } else return false;
    if (r > maxrank || n < off + 8u * r) return false;
    *rank = r;
    for (int i = 0; i < r; ++i) dims[i] = (int64_t)get(d + off + 8 * i, 8);
    return true;
}

int read_vlen_string(const Reader& R, const uint8_t* elem, std::vector<uint8_t>& out)
{
    const uint32_t slen = (uint32_t)get(elem, 4);
    const uint64_t gaddr = get(elem + 4, 8);
    const uint32_t idx = (uint32_t)get(elem + 12, 4);
    out.clear();
    if (slen == 0 || gaddr == 0 || gaddr == UNDEF) return CGB_OK;
    if (!R.ok(gaddr, 16) || memcmp(R.at(gaddr), "GCOL", 4) != 0) CGB_H5_IOFAIL("cgb_h5: bad global heap address");
    const uint64_t size = get(R.at(gaddr) + 8, 8);
    if (!R.ok(gaddr, size)) CGB_H5_IOFAIL("cgb_h5: global heap collection exceeds the file");
    uint64_t o = 16;
    while (o + 16 <= size) {
        const uint8_t* q = R.at(gaddr) + o;
        const uint32_t oi = (uint32_t)get(q, 2);
        const uint64_t osz = get(q + 8, 8);
        if (oi == 0) break;
        if (oi == idx) {
            if (o + 16 + osz > size) CGB_H5_IOFAIL("cgb_h5: global heap object exceeds its collection");
            out.assign(q + 16, q + 16 + (slen < osz ? slen : osz));
            return CGB_OK;
        }
        o += 16 + (osz + 7) / 8 * 8;
    }
    CGB_H5_IOFAIL("cgb_h5: global heap object %u not found", idx);
}

int decode_attr(const Reader& R, const uint8_t* d, size_t n, Attr& a)
{
    if (n < 8) CGB_H5_IOFAIL("cgb_h5: truncated attribute message");
    const int ver = d[0];
    if (ver < 1 || ver > 3) CGB_H5_IOFAIL("cgb_h5: attribute message version %d not supported", ver);
    if (ver >= 2 && (d[1] & 3)) {
        a.dtype = 0;  // shared datatype/dataspace: outside the subset
        return CGB_OK;
    }
    const size_t nsz = get(d + 2, 2), tsz = get(d + 4, 2), ssz = get(d + 6, 2);
    size_t off = ver == 3 ? 9 : 8;
    auto adv = [&](size_t s) { return ver == 1 ? (s + 7) / 8 * 8 : s; };
    if (off + adv(nsz) + adv(tsz) + adv(ssz) > n || nsz == 0) CGB_H5_IOFAIL("cgb_h5: truncated attribute message");
    a.name.assign((const char*)d + off, strnlen((const char*)d + off, nsz));
    off += adv(nsz);
    const uint8_t* dt = d + off;
    off += adv(tsz);
    const uint8_t* sp = d + off;
    off += adv(ssz);
    size_t fixed_len = 0;
    a.dtype = decode_dtype(dt, tsz, &fixed_len);
    int rank = 0;
    int64_t dims[4] = {0, 0, 0, 0};
    if (!decode_space(sp, ssz, &rank, dims, 1)) a.dtype = 0;
    if (a.dtype == 0) return CGB_OK;
    a.count = rank == 0 ? -1 : dims[0];
    const size_t nelem = rank == 0 ? 1 : (size_t)dims[0];
    if (a.dtype == CGB_H5_STR) {
        if (nelem != 1) {
            a.dtype = 0;  // arrays of strings: outside the subset
            return CGB_OK;
        }
        a.count = -1;
        if (fixed_len) {
            if (off + fixed_len > n) CGB_H5_IOFAIL("cgb_h5: truncated attribute data");
            a.data.assign(d + off, d + off + strnlen((const char*)d + off, fixed_len));
        } else {
            if (off + 16 > n) CGB_H5_IOFAIL("cgb_h5: truncated attribute data");
            CGB_H5_CHECK(read_vlen_string(R, d + off, a.data));
        }
        return CGB_OK;
    }
    const size_t bytes = nelem * dtype_size(a.dtype);
    if (off + bytes > n) CGB_H5_IOFAIL("cgb_h5: truncated attribute data");
    a.data.assign(d + off, d + off + bytes);
    return CGB_OK;
}

struct HeaderInfo {
    bool has_stab = false;
    uint64_t btree = UNDEF, heap = UNDEF;
};

int parse_messages(const Reader& R, uint64_t addr, uint64_t size, int* remaining, Node& n, HeaderInfo& hi, int depth)
{
    if (depth > 64) CGB_H5_IOFAIL("cgb_h5: object header continuation chain too long");
    if (!R.ok(addr, size)) CGB_H5_IOFAIL("cgb_h5: object header block exceeds the file");
    const uint8_t* p = R.at(addr);
    uint64_t o = 0;
    while (*remaining > 0 && o + 8 <= size) {
        const uint16_t type = (uint16_t)get(p + o, 2);
        const uint16_t msz = (uint16_t)get(p + o + 2, 2);
        const uint8_t flags = p[o + 4];
        const uint8_t* d = p + o + 8;
        if (o + 8 + msz > size) CGB_H5_IOFAIL("cgb_h5: object header message exceeds its block");
        --*remaining;
        o += 8 + msz;
        const bool shared = flags & 2;
        switch (type) {
        case 0x0001:
            if (shared || !decode_space(d, msz, &n.rank, n.dims, 4)) n.dtype = -1;
            break;
        case 0x0003: {
            size_t fl = 0;
            const int dt = shared ? 0 : decode_dtype(d, msz, &fl);
            if (n.dtype != -1) n.dtype = (dt == 0 || dt == CGB_H5_STR) ? -1 : dt;
            n.is_group = false;
            break;
        }
        case 0x0008: {
            n.is_group = false;
            if (msz < 2) break;
            const int ver = d[0];
            if (ver == 1 || ver == 2) {
                const int rk = d[1], cls = d[2];
                if (cls == 1 && msz >= 16) {
                    n.data_addr = get(d + 8, 8);
                } else if (cls == 0 && msz >= 8 + 4u * rk + 4) {
                    const uint32_t csz = (uint32_t)get(d + 8 + 4 * rk, 4);
                    if (8 + 4u * rk + 4 + csz <= msz) n.compact.assign(d + 12 + 4 * rk, d + 12 + 4 * rk + csz);
                    n.data_addr = 0;
                } else n.dtype = -1;
            } else if (ver == 3 || ver == 4) {
                const int cls = d[1];
                if (cls == 1 && msz >= 18) {
                    n.data_addr = get(d + 2, 8);
                } else if (cls == 0 && msz >= 4) {
                    const uint32_t csz = (uint32_t)get(d + 2, 2);
                    if (4u + csz <= msz) n.compact.assign(d + 4, d + 4 + csz);
                    n.data_addr = 0;
                } else n.dtype = -1;  // chunked / virtual: outside the subset
            } else n.dtype = -1;
            break;
        }
        case 0x000C: {
            Attr a;
            if (!shared) {
                CGB_H5_CHECK(decode_attr(R, d, msz, a));
                n.attrs.push_back(std::move(a));
            }
            break;
        }
        case 0x0010:
            if (msz >= 16) CGB_H5_CHECK(parse_messages(R, get(d, 8), get(d + 8, 8), remaining, n, hi, depth + 1));
            break;
        case 0x0011:
            if (msz >= 16) {
                hi.has_stab = true;
                hi.btree = get(d, 8);
                hi.heap = get(d + 8, 8);
            }
            break;
        default: break;
        }
    }
    return CGB_OK;
}

int load_object(const Reader& R, uint64_t addr, Node& n, int depth);

int load_btree(const Reader& R, uint64_t addr, uint64_t heap_data, uint64_t heap_size, Node& group, int depth)
{
    if (depth > 32) CGB_H5_IOFAIL("cgb_h5: group B-tree too deep");
    if (!R.ok(addr, 24)) CGB_H5_IOFAIL("cgb_h5: B-tree node exceeds the file");
    const uint8_t* p = R.at(addr);
    if (memcmp(p, "TREE", 4) != 0 || p[4] != 0) CGB_H5_IOFAIL("cgb_h5: bad group B-tree node");
    const int level = p[5];
    const int used = (int)get(p + 6, 2);
    if (!R.ok(addr, 24 + 16ull * used + 8)) CGB_H5_IOFAIL("cgb_h5: B-tree node exceeds the file");
    for (int i = 0; i < used; ++i) {
        const uint64_t child = get(p + 24 + 16 * i + 8, 8);
        if (level > 0) {
            CGB_H5_CHECK(load_btree(R, child, heap_data, heap_size, group, depth + 1));
            continue;
        }
        if (!R.ok(child, 8)) CGB_H5_IOFAIL("cgb_h5: symbol-table node exceeds the file");
        const uint8_t* s = R.at(child);
        if (memcmp(s, "SNOD", 4) != 0) CGB_H5_IOFAIL("cgb_h5: bad symbol-table node");
        const int ns = (int)get(s + 6, 2);
        if (!R.ok(child, 8 + 40ull * ns)) CGB_H5_IOFAIL("cgb_h5: symbol-table node exceeds the file");
        for (int e = 0; e < ns; ++e) {
            const uint8_t* q = s + 8 + 40 * e;
            const uint64_t noff = get(q, 8);
            const uint64_t oaddr = get(q + 8, 8);
            if (noff >= heap_size) CGB_H5_IOFAIL("cgb_h5: link name outside the local heap");
            const char* nm = (const char*)R.at(heap_data) + noff;
            std::string name(nm, strnlen(nm, heap_size - noff));
            std::unique_ptr<Node> c(new Node);
            CGB_H5_CHECK(load_object(R, oaddr, *c, depth + 1));
            group.kids[name] = std::move(c);
        }
    }
    return CGB_OK;
}

int load_object(const Reader& R, uint64_t addr, Node& n, int depth)
{
    if (depth > 32) CGB_H5_IOFAIL("cgb_h5: group nesting too deep");
    if (!R.ok(addr, 16)) CGB_H5_IOFAIL("cgb_h5: object header exceeds the file");
    const uint8_t* p = R.at(addr);
    if (p[0] != 1) CGB_H5_IOFAIL("cgb_h5: object header version %d not supported (file written with libver='latest'?)", p[0]);
    int nmsgs = (int)get(p + 2, 2);
    const uint64_t hsize = get(p + 8, 4);
    HeaderInfo hi;
    n.is_group = true;
    CGB_H5_CHECK(parse_messages(R, addr + 16, hsize, &nmsgs, n, hi, 0));
    if (hi.has_stab) {
        n.is_group = true;
        if (!R.ok(hi.heap, 32) || memcmp(R.at(hi.heap), "HEAP", 4) != 0) CGB_H5_IOFAIL("cgb_h5: bad local heap");
        const uint64_t seg_size = get(R.at(hi.heap) + 8, 8);
        const uint64_t seg_addr = get(R.at(hi.heap) + 24, 8);
        if (!R.ok(seg_addr, seg_size)) CGB_H5_IOFAIL("cgb_h5: local heap data exceeds the file");
        CGB_H5_CHECK(load_btree(R, hi.btree, seg_addr, seg_size, n, depth));
    } else if (n.is_group) {
        CGB_H5_IOFAIL("cgb_h5: object is neither a symbol-table group nor a dataset (new-style group?)");
    } else if (n.dtype > 0) {
        uint64_t count = 1;
        for (int i = 0; i < n.rank; ++i) count *= (uint64_t)n.dims[i];
        n.data_bytes = count * dtype_size(n.dtype);
        if (!n.compact.empty() || n.data_addr == 0) {
            if (n.compact.size() < n.data_bytes) n.dtype = -1;
        } else if (n.data_bytes && !R.ok(n.data_addr, n.data_bytes)) {
            if (n.data_addr == UNDEF) n.dtype = -1;  // never written
            else CGB_H5_IOFAIL("cgb_h5: dataset data exceeds the file");
        }
    }
    return CGB_OK;
}

Attr* find_attr(Node* n, const char* name)
{
    for (Attr& a : n->attrs)
        if (a.name == name) return &a;
    return nullptr;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int cgb_h5_create(const char* path, int overwrite, cgb_h5_file** out)
{
    CGB_REQUIRE(path && out, "cgb_h5_create: null argument");
    const int fd = open(path, O_CREAT | O_WRONLY | (overwrite ? O_TRUNC : O_EXCL), 0644);
    if (fd < 0) CGB_H5_IOFAIL("Cannot create %s: %s", path, strerror(errno));
    cgb_h5_file* f = new cgb_h5_file;
    f->fd = fd;
    f->writing = true;
    f->path = path;
    f->cursor = 4096;  // superblock page; datasets start page-aligned
    *out = f;
    return CGB_OK;
}

int cgb_h5_create_group(cgb_h5_file* f, const char* path)
{
    CGB_REQUIRE(f && f->writing && path, "cgb_h5_create_group: file is not open for writing");
    Node* parent;
    std::string leaf;
    CGB_H5_CHECK(find_parent(f, path, &parent, &leaf));
    parent->kids[leaf] = std::unique_ptr<Node>(new Node);
    return CGB_OK;
}

int cgb_h5_write_dataset(cgb_h5_file* f, const char* path, int dtype, int rank, const int64_t* dims,
                         const void* host_data)
{
    Node* n;
    CGB_H5_CHECK(new_dataset(f, path, dtype, rank, dims, &n));
    CGB_REQUIRE(host_data || n->data_bytes == 0, "cgb_h5_write_dataset: null data");
    return pwrite_all(f->fd, host_data, n->data_bytes, n->data_addr);
}

int cgb_h5_write_dataset_device(cgb_h5_file* f, const char* path, int out_dtype, const void* d_src, int src_dtype,
                                int64_t n, int ncomp, int64_t ld, double scale, void* stream)
{
    CGB_REQUIRE(src_dtype == CGB_H5_F32 || src_dtype == CGB_H5_F64, "cgb_h5_write_dataset_device: source must be f32/f64");
    CGB_REQUIRE(out_dtype == CGB_H5_F32 || out_dtype == CGB_H5_F64, "cgb_h5_write_dataset_device: output must be f32/f64");
    CGB_REQUIRE(n >= 0 && ncomp >= 1 && ncomp <= 4 && (ncomp == 1 || ld >= n), "cgb_h5_write_dataset_device: bad shape");
    const int64_t dims[2] = {n, ncomp};
    Node* node;
    CGB_H5_CHECK(new_dataset(f, path, out_dtype, ncomp == 1 ? 1 : 2, dims, &node));
    if (n == 0) return CGB_OK;
    CGB_REQUIRE(d_src, "cgb_h5_write_dataset_device: null source");
    cudaStream_t st = (cudaStream_t)stream;
    CGB_H5_CHECK(ensure_staging(f));
    const size_t row = (size_t)ncomp * dtype_size(out_dtype);
    const int64_t per_chunk = (int64_t)(STAGE_BYTES / row);
    const bool plain = ncomp == 1 && src_dtype == out_dtype && scale == 1.0;
    const int64_t nchunks = (n + per_chunk - 1) / per_chunk;
    for (int64_t c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {
            const int slot = (int)(c & 1);
            const int64_t first = c * per_chunk;
            const int64_t m = n - first < per_chunk ? n - first : per_chunk;
            if (plain) {
                CGB_CUDA_CHECK(cudaMemcpyAsync(f->pinned[slot], (const char*)d_src + first * row, m * row,
                                               cudaMemcpyDeviceToHost, st));
            } else {
                if (src_dtype == CGB_H5_F64 && out_dtype == CGB_H5_F64)
                    launch_pack<double, double>(d_src, ld, first, m, ncomp, scale, f->d_stage[slot], st);
                else if (src_dtype == CGB_H5_F64)
                    launch_pack<double, float>(d_src, ld, first, m, ncomp, scale, f->d_stage[slot], st);
                else if (out_dtype == CGB_H5_F64)
                    launch_pack<float, double>(d_src, ld, first, m, ncomp, scale, f->d_stage[slot], st);
                else
                    launch_pack<float, float>(d_src, ld, first, m, ncomp, scale, f->d_stage[slot], st);
                CGB_CUDA_CHECK(cudaGetLastError());
                CGB_CUDA_CHECK(cudaMemcpyAsync(f->pinned[slot], f->d_stage[slot], m * row, cudaMemcpyDeviceToHost, st));
            }
            CGB_CUDA_CHECK(cudaEventRecord(f->ev[slot], st));
        }
        if (c > 0) {  // file write of the previous chunk overlaps the copy just queued
            const int slot = (int)((c - 1) & 1);
            const int64_t first = (c - 1) * per_chunk;
            const int64_t m = n - first < per_chunk ? n - first : per_chunk;
            CGB_CUDA_CHECK(cudaEventSynchronize(f->ev[slot]));
            CGB_H5_CHECK(pwrite_all(f->fd, f->pinned[slot], m * row, node->data_addr + first * row));
        }
    }
    return CGB_OK;
}

int cgb_h5_set_attr(cgb_h5_file* f, const char* obj_path, const char* name, int dtype, int64_t count,
                    const void* data)
{
    CGB_REQUIRE(f && f->writing && obj_path && name && *name, "cgb_h5_set_attr: file is not open for writing");
    Node* n = find_node(f, obj_path);
    CGB_REQUIRE(n, "cgb_h5_set_attr: object '%s' does not exist", obj_path);
    Attr a;
    a.name = name;
    a.dtype = dtype;
    if (dtype == CGB_H5_STR) {
        CGB_REQUIRE(data, "cgb_h5_set_attr: null string");
        const char* s = (const char*)data;
        a.count = -1;
        a.data.assign(s, s + strlen(s));
    } else {
        CGB_REQUIRE(dtype_size(dtype) > 0, "cgb_h5_set_attr: dtype %d not supported", dtype);
        CGB_REQUIRE(count >= -1 && (data || count == 0), "cgb_h5_set_attr: bad count/data");
        a.count = count;
        const size_t bytes = (size_t)(count < 0 ? 1 : count) * dtype_size(dtype);
        a.data.assign((const uint8_t*)data, (const uint8_t*)data + bytes);
    }
    if (Attr* old = find_attr(n, name)) *old = std::move(a);
    else n->attrs.push_back(std::move(a));
    return CGB_OK;
}

int cgb_h5_close(cgb_h5_file* f)
{
    if (!f) return CGB_OK;
    int rc = CGB_OK;
    if (f->writing) {
        rc = finish_file(f);
        if (rc == CGB_OK && fsync(f->fd) != 0 && errno != EINVAL && errno != EROFS) {
            cgb_set_error("cgb_h5_close: fsync failed: %s", strerror(errno));
            rc = CGB_E_IO;
        }
    }
    release(f);
    return rc;
}

int cgb_h5_open(const char* path, cgb_h5_file** out)
{
    CGB_REQUIRE(path && out, "cgb_h5_open: null argument");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) CGB_H5_IOFAIL("Cannot open %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 96) {
        close(fd);
        CGB_H5_IOFAIL("%s is not an HDF5 file (too small)", path);
    }
    void* mp = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (mp == MAP_FAILED) {
        close(fd);
        CGB_H5_IOFAIL("cgb_h5_open: mmap failed: %s", strerror(errno));
    }
    cgb_h5_file* f = new cgb_h5_file;
    f->fd = fd;
    f->path = path;
    f->map = (const uint8_t*)mp;
    f->map_len = (size_t)sb.st_size;
    // the superblock sits at 0 or at a power-of-two offset >= 512 (user block)
    uint64_t sboff = UNDEF;
    for (uint64_t o = 0; o + 96 <= f->map_len; o = o ? o * 2 : 512) {
        if (memcmp(f->map + o, H5_SIG, 8) == 0) {
            sboff = o;
            break;
        }
    }
    int rc = CGB_OK;
    if (sboff == UNDEF) {
        cgb_set_error("%s is not an HDF5 file (no signature)", path);
        rc = CGB_E_IO;
    } else {
        const uint8_t* s = f->map + sboff;
        const int ver = s[8];
        if (ver > 1 || s[13] != 8 || s[14] != 8) {
            cgb_set_error("cgb_h5_open: superblock version %d / %d-byte offsets not supported "
                          "(only the default 'earliest' HDF5 layout is)", ver, s[13]);
            rc = CGB_E_IO;
        } else {
            const uint8_t* a = s + (ver == 0 ? 24 : 28);
            f->base = get(a, 8);
            Reader R{f->map, f->map_len, f->base};
            const uint64_t root_hdr = get(a + 32 + 8, 8);
            rc = load_object(R, root_hdr, f->root, 0);
        }
    }
    if (rc != CGB_OK) {
        release(f);
        return rc;
    }
    *out = f;
    return CGB_OK;
}

int cgb_h5_num_children(cgb_h5_file* f, const char* group_path, int* n)
{
    CGB_REQUIRE(f && group_path && n, "cgb_h5_num_children: null argument");
    Node* g = find_node(f, group_path);
    CGB_REQUIRE(g && g->is_group, "cgb_h5: no group '%s'", group_path);
    *n = (int)g->kids.size();
    return CGB_OK;
}

int cgb_h5_child_name(cgb_h5_file* f, const char* group_path, int i, char* buf, int buflen, int* is_group)
{
    CGB_REQUIRE(f && group_path && buf && buflen > 0, "cgb_h5_child_name: null argument");
    Node* g = find_node(f, group_path);
    CGB_REQUIRE(g && g->is_group, "cgb_h5: no group '%s'", group_path);
    CGB_REQUIRE(i >= 0 && i < (int)g->kids.size(), "cgb_h5_child_name: index out of range");
    auto it = g->kids.begin();
    std::advance(it, i);
    CGB_REQUIRE((int)it->first.size() < buflen, "cgb_h5_child_name: buffer too small");
    memcpy(buf, it->first.c_str(), it->first.size() + 1);
    if (is_group) *is_group = it->second->is_group ? 1 : 0;
    return CGB_OK;
}

int cgb_h5_dataset_info(cgb_h5_file* f, const char* path, int* dtype, int* rank, int64_t* dims4,
                        int64_t* file_offset)
{
    CGB_REQUIRE(f && path, "cgb_h5_dataset_info: null argument");
    Node* n = find_node(f, path);
    CGB_REQUIRE(n && !n->is_group, "cgb_h5: no dataset '%s'", path);
    CGB_REQUIRE(n->dtype > 0, "cgb_h5: dataset '%s' has a type or layout outside the supported subset", path);
    if (dtype) *dtype = n->dtype;
    if (rank) *rank = n->rank;
    if (dims4)
        for (int i = 0; i < 4; ++i) dims4[i] = n->dims[i];
    if (file_offset) *file_offset = n->compact.empty() && n->data_bytes ? (int64_t)(n->data_addr + f->base) : -1;
    return CGB_OK;
}

int cgb_h5_read_dataset(cgb_h5_file* f, const char* path, void* host_out, size_t bytes)
{
    CGB_REQUIRE(f && path && !f->writing, "cgb_h5_read_dataset: file is not open for reading");
    Node* n = find_node(f, path);
    CGB_REQUIRE(n && !n->is_group, "cgb_h5: no dataset '%s'", path);
    CGB_REQUIRE(n->dtype > 0, "cgb_h5: dataset '%s' has a type or layout outside the supported subset", path);
    CGB_REQUIRE(bytes == n->data_bytes, "cgb_h5_read_dataset: buffer is %zu bytes, dataset is %llu", bytes,
                (unsigned long long)n->data_bytes);
    if (bytes == 0) return CGB_OK;
    CGB_REQUIRE(host_out, "cgb_h5_read_dataset: null buffer");
    if (!n->compact.empty()) memcpy(host_out, n->compact.data(), bytes);
    else memcpy(host_out, f->map + f->base + n->data_addr, bytes);
    return CGB_OK;
}

int cgb_h5_read_dataset_device(cgb_h5_file* f, const char* path, void* d_dst, int dst_dtype, int64_t ld,
                               double scale, void* stream)
{
    CGB_REQUIRE(f && path && !f->writing, "cgb_h5_read_dataset_device: file is not open for reading");
    Node* n = find_node(f, path);
    CGB_REQUIRE(n && !n->is_group, "cgb_h5: no dataset '%s'", path);
    CGB_REQUIRE(n->dtype == CGB_H5_F32 || n->dtype == CGB_H5_F64,
                "cgb_h5_read_dataset_device: dataset '%s' is not f32/f64", path);
    CGB_REQUIRE(dst_dtype == CGB_H5_F32 || dst_dtype == CGB_H5_F64, "cgb_h5_read_dataset_device: output must be f32/f64");
    CGB_REQUIRE(n->rank == 1 || n->rank == 2, "cgb_h5_read_dataset_device: dataset '%s' must have rank 1 or 2", path);
    const int64_t rows = n->dims[0];
    const int ncomp = n->rank == 2 ? (int)n->dims[1] : 1;
    CGB_REQUIRE(ncomp >= 1 && ncomp <= 4 && (ncomp == 1 || ld >= rows), "cgb_h5_read_dataset_device: bad shape");
    if (rows == 0) return CGB_OK;
    CGB_REQUIRE(d_dst, "cgb_h5_read_dataset_device: null destination");
    const uint8_t* src = n->compact.empty() ? f->map + f->base + n->data_addr : n->compact.data();
    cudaStream_t st = (cudaStream_t)stream;
    CGB_H5_CHECK(ensure_staging(f));
    const size_t row = (size_t)ncomp * dtype_size(n->dtype);
    const int64_t per_chunk = (int64_t)(STAGE_BYTES / row);
    const bool plain = ncomp == 1 && n->dtype == dst_dtype && scale == 1.0;
    const int64_t nchunks = (rows + per_chunk - 1) / per_chunk;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int slot = (int)(c & 1);
        const int64_t first = c * per_chunk;
        const int64_t m = rows - first < per_chunk ? rows - first : per_chunk;
        // the slot's previous upload (chunk c - 2) must have left the pinned buffer; the host copy
        // out of the page cache then overlaps the upload + unpack of chunk c - 1
        if (c >= 2) CGB_CUDA_CHECK(cudaEventSynchronize(f->ev[slot]));
        memcpy(f->pinned[slot], src + first * row, m * row);
        if (plain) {
            CGB_CUDA_CHECK(cudaMemcpyAsync((char*)d_dst + first * row, f->pinned[slot], m * row,
                                           cudaMemcpyHostToDevice, st));
        } else {
            CGB_CUDA_CHECK(cudaMemcpyAsync(f->d_stage[slot], f->pinned[slot], m * row, cudaMemcpyHostToDevice, st));
            if (n->dtype == CGB_H5_F64 && dst_dtype == CGB_H5_F64)
                launch_unpack<double, double>(f->d_stage[slot], ld, first, m, ncomp, scale, d_dst, st);
            else if (n->dtype == CGB_H5_F64)
                launch_unpack<double, float>(f->d_stage[slot], ld, first, m, ncomp, scale, d_dst, st);
            else if (dst_dtype == CGB_H5_F64)
                launch_unpack<float, double>(f->d_stage[slot], ld, first, m, ncomp, scale, d_dst, st);
            else
                launch_unpack<float, float>(f->d_stage[slot], ld, first, m, ncomp, scale, d_dst, st);
            CGB_CUDA_CHECK(cudaGetLastError());
        }
        CGB_CUDA_CHECK(cudaEventRecord(f->ev[slot], st));
    }
    CGB_CUDA_CHECK(cudaStreamSynchronize(st));  // the staging slots belong to the handle
    return CGB_OK;
}

int cgb_h5_num_attrs(cgb_h5_file* f, const char* obj_path, int* n)
{
    CGB_REQUIRE(f && obj_path && n, "cgb_h5_num_attrs: null argument");
    Node* o = find_node(f, obj_path);
    CGB_REQUIRE(o, "cgb_h5: no object '%s'", obj_path);
    *n = (int)o->attrs.size();
    return CGB_OK;
}

int cgb_h5_attr_name(cgb_h5_file* f, const char* obj_path, int i, char* buf, int buflen)
{
    CGB_REQUIRE(f && obj_path && buf && buflen > 0, "cgb_h5_attr_name: null argument");
    Node* o = find_node(f, obj_path);
    CGB_REQUIRE(o, "cgb_h5: no object '%s'", obj_path);
    CGB_REQUIRE(i >= 0 && i < (int)o->attrs.size(), "cgb_h5_attr_name: index out of range");
    CGB_REQUIRE((int)o->attrs[i].name.size() < buflen, "cgb_h5_attr_name: buffer too small");
    memcpy(buf, o->attrs[i].name.c_str(), o->attrs[i].name.size() + 1);
    return CGB_OK;
}

int cgb_h5_attr_info(cgb_h5_file* f, const char* obj_path, const char* name, int* dtype, int64_t* count,
                     size_t* bytes)
{
    CGB_REQUIRE(f && obj_path && name, "cgb_h5_attr_info: null argument");
    Node* o = find_node(f, obj_path);
    CGB_REQUIRE(o, "cgb_h5: no object '%s'", obj_path);
    Attr* a = find_attr(o, name);
    CGB_REQUIRE(a, "cgb_h5: object '%s' has no attribute '%s'", obj_path, name);
    if (dtype) *dtype = a->dtype;
    if (count) *count = a->count;
    if (bytes) *bytes = a->data.size();
    return CGB_OK;
}

int cgb_h5_read_attr(cgb_h5_file* f, const char* obj_path, const char* name, void* out, size_t bytes)
{
    CGB_REQUIRE(f && obj_path && name, "cgb_h5_read_attr: null argument");
    Node* o = find_node(f, obj_path);
    CGB_REQUIRE(o, "cgb_h5: no object '%s'", obj_path);
    Attr* a = find_attr(o, name);
    CGB_REQUIRE(a, "cgb_h5: object '%s' has no attribute '%s'", obj_path, name);
    CGB_REQUIRE(a->dtype > 0, "cgb_h5: attribute '%s' has a type outside the supported subset", name);
    CGB_REQUIRE(bytes == a->data.size(), "cgb_h5_read_attr: buffer is %zu bytes, attribute is %zu", bytes, a->data.size());
    if (bytes) memcpy(out, a->data.data(), bytes);
    return CGB_OK;
}

}  // extern "C"
