// gguf_load.cuh — GGUF -> HBM direct load, host code only.
//
// What it replaces: GgufFile::open (src/gguf/mod.rs:23-40: mmap), GgufReader::read (src/gguf/reader.rs:28-110: header, metadata,
// tensor infos, aligned data offset), ModelLoader::parse_config (src/model/loader.rs:62-300: the {arch}.* keys and their defaults)
// and the load loop that copies every tensor into a Vec before the CUDA path copies it again (loader.rs:1341-1365,
// cuda/gpu_only.rs:426-520).  Here the file is mapped once, parsed in place, and every tensor the engine knows goes from the page
// cache to its HBM allocation through two pinned staging buffers: worker threads fill buffer b while the copy engine drains buffer
// b ^ 1 (cudaMemcpyAsync on the load stream); no tensor-sized host allocation, no f32 embedding table (the engine dequantises the
// one row it needs on the device).  Tensor / expert parallel ranks read the same mapping and stage only their shard.
//
// Format rules mirrored from the reference reader: magic 0x46554747, versions 1-3 (v1: 32-bit counts / lengths / dims), metadata
// value types 0..12, general.alignment (Uint32 / Uint64, default 32), data offset = align_up(end of tensor infos), tensor bytes =
// numel / block_elems * block_bytes, bounds-checked against the mapping (mod.rs:34-42).  The typed getters are as strict as the
// reference's (types.rs:78-97: get_u32 accepts Uint32 only, get_f32 Float32 only); a key of another type reads as absent.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <unordered_map>

namespace b200 {

enum GgufValueType { GV_U8 = 0, GV_I8, GV_U16, GV_I16, GV_U32, GV_I32, GV_F32, GV_BOOL, GV_STRING, GV_ARRAY, GV_U64, GV_I64, GV_F64 };

struct GgufValue {
    uint32_t type = 0;
    uint64_t u = 0;       // integer / bool payload (sign-extended for the signed types)
    double f = 0.0;       // F32 / F64 payload
    std::string s;        // STRING payload
    uint32_t arr_type = 0;
    uint64_t arr_len = 0; // ARRAY: element type and count (elements are skipped, not kept: the path needs only the count)
};

struct GgufTensorInfo {
    std::string name;
    int n_dims = 0;
    uint64_t ne[4] = {1, 1, 1, 1};
    uint32_t type = 0;
    uint64_t offset = 0;   // relative to data_offset
    size_t nbytes = 0;     // 0 when the type is not one the engine implements
};

}  // namespace b200

struct b200_gguf {
    int fd = -1;
    const uint8_t* map = nullptr;
    size_t size = 0;
    uint32_t version = 0;
    uint64_t alignment = 32, data_offset = 0;
    std::unordered_map<std::string, b200::GgufValue> kv;
    std::vector<b200::GgufTensorInfo> tensors;
    std::string arch;
};

namespace b200 {

struct GgufCursor {
    const uint8_t* p;
    size_t size, pos = 0;
    bool ok = true;
    bool need(size_t n) {
        if (!ok || n > size - pos) { ok = false; return false; }
        return true;
    }
    template <typename T>
    T get() {
        T v{};
        if (need(sizeof(T))) { memcpy(&v, p + pos, sizeof(T)); pos += sizeof(T); }
        return v;
    }
    uint64_t len(uint32_t version) { return version == 1 ? (uint64_t)get<uint32_t>() : get<uint64_t>(); }   // reader.rs:230, 279, 319
    std::string str(uint32_t version) {
        const uint64_t n = len(version);
        if (!need(n)) return std::string();
        std::string s(reinterpret_cast<const char*>(p + pos), (size_t)n);
        pos += (size_t)n;
        return s;
    }
};

static size_t gguf_scalar_size(uint32_t t) {
    switch (t) {
        case GV_U8: case GV_I8: case GV_BOOL: return 1;
        case GV_U16: case GV_I16: return 2;
        case GV_U32: case GV_I32: case GV_F32: return 4;
        case GV_U64: case GV_I64: case GV_F64: return 8;
        default: return 0;
    }
}

static bool gguf_read_value(GgufCursor& c, uint32_t version, uint32_t type, GgufValue& v, int depth) {
    v.type = type;
    switch (type) {
        case GV_U8: v.u = c.get<uint8_t>(); break;
        case GV_I8: v.u = (uint64_t)(int64_t)c.get<int8_t>(); break;
        case GV_U16: v.u = c.get<uint16_t>(); break;
        case GV_I16: v.u = (uint64_t)(int64_t)c.get<int16_t>(); break;
        case GV_U32: v.u = c.get<uint32_t>(); break;
        case GV_I32: v.u = (uint64_t)(int64_t)c.get<int32_t>(); break;
        case GV_F32: v.f = c.get<float>(); break;
        case GV_BOOL: v.u = c.get<uint8_t>() != 0; break;
        case GV_STRING: v.s = c.str(version); break;
        case GV_U64: v.u = c.get<uint64_t>(); break;
        case GV_I64: v.u = (uint64_t)c.get<int64_t>(); break;
        case GV_F64: v.f = c.get<double>(); break;
        case GV_ARRAY: {
            v.arr_type = c.get<uint32_t>();
            v.arr_len = c.len(version);
            const size_t es = gguf_scalar_size(v.arr_type);
            if (es) {
                if (v.arr_len > (c.size - c.pos) / es) { c.ok = false; return false; }
                c.pos += (size_t)v.arr_len * es;
            } else if (v.arr_type == GV_STRING) {
                for (uint64_t i = 0; i < v.arr_len && c.ok; i++) {
                    const uint64_t n = c.len(version);
                    if (c.need(n)) c.pos += (size_t)n;
                }
            } else if (v.arr_type == GV_ARRAY && depth < 4) {
                for (uint64_t i = 0; i < v.arr_len && c.ok; i++) {
                    GgufValue inner;
                    gguf_read_value(c, version, GV_ARRAY, inner, depth + 1);
                }
            } else {
                return false;   // unknown element type
            }
            break;
        }
        default: return false;  // GgufError::InvalidMetadataType
    }
    return c.ok;
}

// strict typed getters (src/gguf/types.rs:78-97)
static bool gguf_get_u32(const b200_gguf* g, const std::string& key, uint32_t& out) {
    auto it = g->kv.find(key);
    if (it == g->kv.end() || it->second.type != GV_U32) return false;
    out = (uint32_t)it->second.u;
    return true;
}
static bool gguf_get_f32(const b200_gguf* g, const std::string& key, float& out) {
    auto it = g->kv.find(key);
    if (it == g->kv.end() || it->second.type != GV_F32) return false;
    out = (float)it->second.f;
    return true;
}
static const GgufTensorInfo* gguf_find_tensor(const b200_gguf* g, const char* name) {
    for (const GgufTensorInfo& t : g->tensors)
        if (t.name == name) return &t;
    return nullptr;
}

static void gguf_release(b200_gguf* g) {
    if (!g) return;
    if (g->map && g->map != (const uint8_t*)MAP_FAILED) munmap((void*)g->map, g->size);
    if (g->fd >= 0) close(g->fd);
    delete g;
}

// fills `dst` from `src` with `threads` host threads (page-cache reads of a mapped file are page faults + memcpy: one thread
// moves ~3-6 GB/s, a PCIe 5 x16 link takes ~55)
static void gguf_parallel_copy(uint8_t* dst, const uint8_t* src, size_t n, int threads) {
    if (threads <= 1 || n < ((size_t)4 << 20)) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = ((n + threads - 1) / threads + 4095) & ~(size_t)4095;
    for (int t = 1; t < threads; t++) {
        const size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        if (lo < hi) th.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(n, per));
    for (std::thread& t : th) t.join();
}

}  // namespace b200
