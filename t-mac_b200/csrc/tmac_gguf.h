// tmac_gguf.h -- minimal GGUF (v2 / v3, little endian) container reader: metadata key/values and the tensor directory,
// tensor data mapped read-only.  Host side, no CUDA types.
//
// Why: the reference pipeline's product is a .gguf file whose quantised linears are `I1..I4` tensors (permuted weights ||
// fp32 scales, python/t_mac/model_utils.py:271; convert_hf_to_gguf.py:536-588) or stock Q4_0 / TQ1_0 / TQ2_0 tensors that
// llama.cpp hands to ggml_tmac_transform_tensor at load time (src/llama.cpp:5216).  With this reader the library can take
// such a file directly (tmac_b200_gguf_*), without llama.cpp in the loop.
//
// Layout (gguf-py/gguf/gguf_writer.py, gguf_reader.py of the vendored llama.cpp): magic "GGUF", u32 version, u64 tensor
// count, u64 kv count; kv = string key, u32 type, value; tensor info = string name, u32 n_dims, u64 dims[n_dims] (ne0
// first), u32 ggml type, u64 offset into the data section; data section starts at the next multiple of
// `general.alignment` (default 32).  A tensor's byte size is taken from the directory (next offset / end of file), not from
// type x shape: I-type tensors carry their scales behind the packed weights.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace tmac_b200 {

struct GgufTensor {
    std::string name;
    int type = 0, n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    uint64_t offset = 0, nbytes = 0;     // offset from the start of the data section
};

struct GgufValue {
    int type = -1;                        // gguf value type of a scalar / string; arrays: 9
    int elem_type = -1;                   // arrays
    uint64_t count = 0;                   // arrays
    int64_t i = 0;
    double f = 0.0;
    std::string s;
};

struct GgufFile {
    int fd = -1;
    const uint8_t *base = nullptr;
    size_t size = 0;
    uint32_t version = 0;
    uint64_t alignment = 32, data_start = 0;
    std::vector<GgufTensor> tensors;
    std::map<std::string, GgufValue> meta;
    std::string error;

    ~GgufFile() { close(); }
    void close() {
        if (base) munmap((void *)base, size);
        if (fd >= 0) ::close(fd);
        base = nullptr; fd = -1; size = 0;
    }

    struct Cursor {
        const uint8_t *p, *end;
        bool ok = true;
        template <class T> T get() {
            T v{};
            if (!ok || (size_t)(end - p) < sizeof(T)) { ok = false; return v; }
            std::memcpy(&v, p, sizeof(T)); p += sizeof(T);
            return v;
        }
        std::string str() {
            const uint64_t n = get<uint64_t>();
            if (!ok || n > (uint64_t)(end - p)) { ok = false; return std::string(); }
            std::string s((const char *)p, (size_t)n); p += n;
            return s;
        }
        void skip(uint64_t n) { if (!ok || n > (uint64_t)(end - p)) ok = false; else p += n; }
    };

    static int scalar_size(int t) {
        switch (t) { case 0: case 1: case 7: return 1; case 2: case 3: return 2; case 4: case 5: case 6: return 4; case 10: case 11: case 12: return 8; }
        return 0;
    }
    static bool read_scalar(Cursor &c, int t, GgufValue *v) {
        switch (t) {
            case 0: v->i = c.get<uint8_t>(); break;
            case 1: v->i = c.get<int8_t>(); break;
            case 2: v->i = c.get<uint16_t>(); break;
            case 3: v->i = c.get<int16_t>(); break;
            case 4: v->i = c.get<uint32_t>(); break;
            case 5: v->i = c.get<int32_t>(); break;
            case 6: v->f = c.get<float>(); v->i = (int64_t)v->f; break;
            case 7: v->i = c.get<uint8_t>() != 0; break;
            case 10: v->i = (int64_t)c.get<uint64_t>(); break;
            case 11: v->i = c.get<int64_t>(); break;
            case 12: v->f = c.get<double>(); v->i = (int64_t)v->f; break;
            default: return false;
        }
        if (t != 6 && t != 12) v->f = (double)v->i;
        return c.ok;
    }

    bool open(const char *path) {
        close();
        tensors.clear(); meta.clear(); error.clear();
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { error = std::string("cannot open ") + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 24) { error = "not a GGUF file (too short)"; close(); return false; }
        size = (size_t)st.st_size;
        void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { error = "mmap failed"; base = nullptr; close(); return false; }
        base = (const uint8_t *)m;
        Cursor c{base, base + size};
        if (c.get<uint32_t>() != 0x46554747u) { error = "not a GGUF file (bad magic)"; close(); return false; }
        version = c.get<uint32_t>();
        if (version < 2 || version > 3) { error = "unsupported GGUF version " + std::to_string(version); close(); return false; }
        const uint64_t nt = c.get<uint64_t>(), nkv = c.get<uint64_t>();
        if (!c.ok || nt > (1u << 24) || nkv > (1u << 24)) { error = "corrupt GGUF header"; close(); return false; }
        for (uint64_t k = 0; k < nkv; ++k) {
            const std::string key = c.str();
            GgufValue v;
            v.type = (int)c.get<uint32_t>();
            if (v.type == 8) v.s = c.str();
            else if (v.type == 9) {
                v.elem_type = (int)c.get<uint32_t>();
                v.count = c.get<uint64_t>();
                if (v.elem_type == 8) { for (uint64_t e = 0; e < v.count && c.ok; ++e) c.str(); }
                else if (scalar_size(v.elem_type)) c.skip(v.count * (uint64_t)scalar_size(v.elem_type));
                else c.ok = false;
            } else if (!read_scalar(c, v.type, &v)) c.ok = false;
            if (!c.ok) { error = "corrupt GGUF metadata at key '" + key + "'"; close(); return false; }
            meta[key] = v;
        }
        auto al = meta.find("general.alignment");
        if (al != meta.end() && al->second.i > 0) alignment = (uint64_t)al->second.i;
        tensors.resize((size_t)nt);
        for (uint64_t t = 0; t < nt; ++t) {
            GgufTensor &T = tensors[(size_t)t];
            T.name = c.str();
            T.n_dims = (int)c.get<uint32_t>();
            if (!c.ok || T.n_dims < 0 || T.n_dims > 4) { error = "corrupt GGUF tensor directory"; close(); return false; }
            for (int d = 0; d < T.n_dims; ++d) T.ne[d] = (int64_t)c.get<uint64_t>();
            T.type = (int)c.get<uint32_t>();
            T.offset = c.get<uint64_t>();
            if (!c.ok) { error = "corrupt GGUF tensor directory"; close(); return false; }
        }
        const uint64_t pos = (uint64_t)(c.p - base);
        data_start = (pos + alignment - 1) / alignment * alignment;
        if (data_start > size && nt) { error = "GGUF data section missing"; close(); return false; }
        // byte sizes from the directory: distance to the next tensor (by offset), last one to the end of the file
        std::vector<size_t> order(tensors.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tensors[a].offset < tensors[b].offset; });
        for (size_t k = 0; k < order.size(); ++k) {
            GgufTensor &T = tensors[order[k]];
            const uint64_t end = (k + 1 < order.size()) ? tensors[order[k + 1]].offset : (uint64_t)size - data_start;
            if (T.offset > end || data_start + end > size) { error = "GGUF tensor '" + T.name + "' lies outside the file"; close(); return false; }
            T.nbytes = end - T.offset;
            const uint64_t need = nominal_bytes(T);
            if (need && T.nbytes < need) { error = "GGUF tensor '" + T.name + "' is truncated"; close(); return false; }
        }
        return true;
    }
    // Minimum byte size implied by type and shape for the types this library meets (0 = unknown type, not checked).
    static uint64_t nominal_bytes(const GgufTensor &T) {
        uint64_t n = 1;
        for (int d = 0; d < T.n_dims; ++d) n *= (uint64_t)T.ne[d];
        switch (T.type) {
            case 0: return n * 4;                     // F32
            case 1: case 30: return n * 2;            // F16, BF16
            case 2: return n / 32 * 18;               // Q4_0
            case 8: return n / 32 * 34;               // Q8_0
            case 34: return n / 256 * 54;             // TQ1_0
            case 35: return n / 256 * 66;             // TQ2_0
            case 36: case 37: case 38: case 39: return n * (uint64_t)(T.type - 35) / 8;   // I1..I4: packed weights (+ scales behind)
        }
        return 0;
    }
    const uint8_t *data(size_t i) const { return base + data_start + tensors[i].offset; }
    int find(const char *name) const {
        for (size_t i = 0; i < tensors.size(); ++i)
            if (tensors[i].name == name) return (int)i;
        return -1;
    }
};

}  // namespace tmac_b200
