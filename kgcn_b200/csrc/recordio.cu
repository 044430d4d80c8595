// Host-side record IO for the block-diagonal ingest and the checkpoint
// reader (row 4): CRC-32C, TFRecord framing, and a purpose-built wire-format reader for the
// tensorflow.Example messages that kgcn/preprocessing/utils.py:178-214 (convert_to_example) writes and
// task_sparse_gcn.py:93-101,153-166 (tf.io.parse_single_example with a feature_spec) reads.
//
// HOST functions over HOST pointers; no CUDA here.  The reference does this work inside TensorFlow's
// C++ runtime (tf.data.TFRecordDataset + the ParseExample op); neither is available, so the formats
// are restated from their published definitions:
//   * TFRecord:  u64le length | u32le masked_crc32c(length bytes) | data | u32le masked_crc32c(data)
//                masked(c) = rotr(c,15) + 0xa282ead8
//   * Example:   message Example { Features features = 1; }
//                message Features { map<string, Feature> feature = 1; }   (map entry: key = 1, value = 2)
//                message Feature { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2;
//                                               Int64List int64_list = 3; } }
//                BytesList { repeated bytes value = 1; }  FloatList { repeated float value = 1 [packed]; }
//                Int64List { repeated int64 value = 1 [packed]; }
//   Both the packed and the unpacked encoding of the repeated scalars are accepted, as any protobuf
//   parser must; when a map key occurs twice the last entry wins (protobuf map semantics).
#include <cstring>

#include "common.cuh"

namespace kgcn {
namespace {

struct Crc32cTables {
    uint32_t t[8][256];
    Crc32cTables() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82f63b78u : 0u);  // reflected Castagnoli
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; ++i)
            for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xffu];
    }
};

uint32_t crc32c(const uint8_t* p, size_t n) {
    static const Crc32cTables T;
    uint32_t c = 0xffffffffu;
    while (n && (reinterpret_cast<uintptr_t>(p) & 7u)) {
        c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xffu];
        --n;
    }
    while (n >= 8) {  // slicing-by-8
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= c;
        c = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^ T.t[4][(w >> 24) & 0xff] ^
            T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^ T.t[1][(w >> 48) & 0xff] ^ T.t[0][(w >> 56) & 0xff];
        p += 8;
        n -= 8;
    }
    while (n--) c = (c >> 8) ^ T.t[0][(c ^ *p++) & 0xffu];
    return c ^ 0xffffffffu;
}

inline uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

inline uint32_t load_u32(const uint8_t* p) {
    return uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24;
}
inline uint64_t load_u64(const uint8_t* p) { return uint64_t(load_u32(p)) | uint64_t(load_u32(p + 4)) << 32; }

// ---- protobuf wire format ----------------------------------------------------------------------
struct Span {
    const uint8_t* p;
    const uint8_t* end;
    bool empty() const { return p >= end; }
};

bool read_varint(Span& s, uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64 && s.p < s.end; shift += 7) {
        const uint8_t b = *s.p++;
        v |= uint64_t(b & 0x7f) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}

bool read_len(Span& s, Span& sub) {
    uint64_t n;
    if (!read_varint(s, n) || n > uint64_t(s.end - s.p)) return false;
    sub = Span{s.p, s.p + n};
    s.p += n;
    return true;
}

bool skip_field(Span& s, uint32_t wire) {
    uint64_t v;
    Span sub;
    switch (wire) {
        case 0: return read_varint(s, v);
        case 1: if (s.end - s.p < 8) return false; s.p += 8; return true;
        case 2: return read_len(s, sub);
        case 5: if (s.end - s.p < 4) return false; s.p += 4; return true;
        default: return false;  // groups are not used by Example
    }
}

// Finds the serialized Feature stored under `key` in one Example; found = false if the key is absent.
bool find_feature(Span ex, const char* key, size_t key_len, Span& feature, bool& found) {
    found = false;
    while (!ex.empty()) {
        uint64_t tag;
        if (!read_varint(ex, tag)) return false;
        if (tag != ((1u << 3) | 2u)) {  // Example.features
            if (!skip_field(ex, tag & 7u)) return false;
            continue;
        }
        Span feats;
        if (!read_len(ex, feats)) return false;
        while (!feats.empty()) {
            if (!read_varint(feats, tag)) return false;
            if (tag != ((1u << 3) | 2u)) {  // Features.feature (map entry)
                if (!skip_field(feats, tag & 7u)) return false;
                continue;
            }
            Span entry;
            if (!read_len(feats, entry)) return false;
            Span k{nullptr, nullptr}, v{nullptr, nullptr};
            while (!entry.empty()) {
                if (!read_varint(entry, tag)) return false;
                if (tag == ((1u << 3) | 2u)) {
                    if (!read_len(entry, k)) return false;
                } else if (tag == ((2u << 3) | 2u)) {
                    if (!read_len(entry, v)) return false;
                } else if (!skip_field(entry, tag & 7u)) {
                    return false;
                }
            }
            if (size_t(k.end - k.p) == key_len && (key_len == 0 || memcmp(k.p, key, key_len) == 0)) {
                feature = v;
                found = true;  // keep scanning: a later duplicate of the key replaces this one
            }
        }
    }
    return true;
}

// Decodes the values of one Feature.  kind: 1 float, 2 int64 (the two kinds of the feature_spec in
// task_sparse_gcn.py:153-166; bytes lists are not used on this path).  A Feature of another kind
// yields zero values plus other_kind = true (tf.io.parse_single_example raises on that).
bool decode_feature(Span f, int kind, uint8_t* out, int64_t capacity, int64_t& n, bool& other_kind) {
    const size_t width = kind == 1 ? 4 : 8;
    while (!f.empty()) {
        uint64_t tag;
        if (!read_varint(f, tag)) return false;
        const uint32_t field = uint32_t(tag >> 3), wire = uint32_t(tag & 7u);
        if (wire != 2 || field < 1 || field > 3) {
            if (!skip_field(f, wire)) return false;
            continue;
        }
        Span list;
        if (!read_len(f, list)) return false;
        if (int(field) - 1 != kind) {
            if (!list.empty()) other_kind = true;
            continue;
        }
        while (!list.empty()) {
            if (!read_varint(list, tag)) return false;
            const uint32_t lf = uint32_t(tag >> 3), lw = uint32_t(tag & 7u);
            if (lf != 1) {
                if (!skip_field(list, lw)) return false;
                continue;
            }
            if (lw == 2) {  // packed
                Span pk;
                if (!read_len(list, pk)) return false;
                if (kind == 1) {
                    if ((pk.end - pk.p) % 4) return false;
                    const int64_t cnt = (pk.end - pk.p) / 4;
                    if (out && n + cnt <= capacity) memcpy(out + size_t(n) * width, pk.p, size_t(cnt) * 4);
                    n += cnt;
                } else {
                    while (!pk.empty()) {
                        uint64_t v;
                        if (!read_varint(pk, v)) return false;
                        if (out && n < capacity) memcpy(out + size_t(n) * width, &v, 8);
                        ++n;
                    }
                }
            } else if (kind == 1 && lw == 5) {  // unpacked float
                if (list.end - list.p < 4) return false;
                if (out && n < capacity) memcpy(out + size_t(n) * width, list.p, 4);
                list.p += 4;
                ++n;
            } else if (kind == 2 && lw == 0) {  // unpacked int64
                uint64_t v;
                if (!read_varint(list, v)) return false;
                if (out && n < capacity) memcpy(out + size_t(n) * width, &v, 8);
                ++n;
            } else {
                return false;
            }
        }
    }
    return true;
}

}  // namespace
}  // namespace kgcn

extern "C" uint32_t kgcn_crc32c(const void* data, size_t n_bytes) {
    if (data == nullptr || n_bytes == 0) return 0u;
    return kgcn::crc32c(static_cast<const uint8_t*>(data), n_bytes);
}

extern "C" uint32_t kgcn_crc32c_masked(const void* data, size_t n_bytes) {
    return kgcn::mask_crc(kgcn_crc32c(data, n_bytes));
}

extern "C" int kgcn_tfrecord_scan(const void* file, size_t n_bytes, int32_t verify_crc, int64_t* rec_off,
                                  int64_t* rec_len, int64_t capacity, int64_t* n_records) {
    KGCN_REQUIRE(n_records != nullptr && (file != nullptr || n_bytes == 0), KGCN_ERR_NULL,
                 "kgcn_tfrecord_scan: NULL pointer");
    KGCN_REQUIRE(capacity == 0 || (rec_off != nullptr && rec_len != nullptr), KGCN_ERR_NULL,
                 "kgcn_tfrecord_scan: NULL output with capacity %lld", (long long)capacity);
    const uint8_t* base = static_cast<const uint8_t*>(file);
    size_t pos = 0;
    int64_t count = 0;
    while (pos < n_bytes) {
        KGCN_REQUIRE(n_bytes - pos >= 12, KGCN_ERR_BAD_SHAPE,
                     "kgcn_tfrecord_scan: truncated record header at byte %zu (record %lld)", pos, (long long)count);
        const uint64_t len = kgcn::load_u64(base + pos);
        if (verify_crc)
            KGCN_REQUIRE(kgcn::mask_crc(kgcn::crc32c(base + pos, 8)) == kgcn::load_u32(base + pos + 8), KGCN_ERR_BAD_SHAPE,
                         "kgcn_tfrecord_scan: corrupted length of record %lld at byte %zu", (long long)count, pos);
        KGCN_REQUIRE(len <= n_bytes - pos - 12 && n_bytes - pos - 12 - len >= 4, KGCN_ERR_BAD_SHAPE,
                     "kgcn_tfrecord_scan: truncated record %lld at byte %zu (length %llu)", (long long)count, pos,
                     (unsigned long long)len);
        const uint8_t* data = base + pos + 12;
        if (verify_crc)
            KGCN_REQUIRE(kgcn::mask_crc(len ? kgcn::crc32c(data, len) : 0u) == kgcn::load_u32(data + len), KGCN_ERR_BAD_SHAPE,
                         "kgcn_tfrecord_scan: corrupted data of record %lld at byte %zu", (long long)count, pos);
        if (count < capacity) {
            rec_off[count] = int64_t(pos + 12);
            rec_len[count] = int64_t(len);
        }
        ++count;
        pos += 12 + len + 4;
    }
    *n_records = count;
    return KGCN_OK;
}

extern "C" int kgcn_tfexample_gather(const void* file, const int64_t* rec_off, const int64_t* rec_len, int64_t n_rec,
                                     const char* key, int32_t kind, void* values, int64_t capacity, int64_t* counts,
                                     int64_t* total) {
    KGCN_REQUIRE(file != nullptr && rec_off != nullptr && rec_len != nullptr && key != nullptr && total != nullptr,
                 KGCN_ERR_NULL, "kgcn_tfexample_gather: NULL pointer");
    KGCN_REQUIRE(kind != 0, KGCN_ERR_UNSUPPORTED, "kgcn_tfexample_gather: bytes lists are not supported (kind 0)");
    KGCN_REQUIRE((kind == 1 || kind == 2) && n_rec >= 0 && capacity >= 0, KGCN_ERR_BAD_SHAPE,
                 "kgcn_tfexample_gather: bad shape (kind %d, n_rec %lld, capacity %lld)", kind, (long long)n_rec,
                 (long long)capacity);
    const uint8_t* base = static_cast<const uint8_t*>(file);
    const size_t key_len = strlen(key);
    const size_t width = kind == 1 ? 4 : 8;
    uint8_t* out = static_cast<uint8_t*>(values);
    int64_t n = 0;
    for (int64_t r = 0; r < n_rec; ++r) {
        KGCN_REQUIRE(rec_off[r] >= 0 && rec_len[r] >= 0, KGCN_ERR_BAD_SHAPE,
                     "kgcn_tfexample_gather: negative offset/length for record %lld", (long long)r);
        kgcn::Span ex{base + rec_off[r], base + rec_off[r] + rec_len[r]}, feat{nullptr, nullptr};
        bool found = false, other = false;
        KGCN_REQUIRE(kgcn::find_feature(ex, key, key_len, feat, found), KGCN_ERR_BAD_SHAPE,
                     "kgcn_tfexample_gather: record %lld is not a valid Example message", (long long)r);
        int64_t got = 0;
        if (found) {
            // decode into the remaining room; the counters keep running when the room is exhausted
            int64_t local = 0;
            uint8_t* dst = (out != nullptr && n <= capacity) ? out + size_t(n) * width : nullptr;
            KGCN_REQUIRE(kgcn::decode_feature(feat, kind, dst, dst ? capacity - n : 0, local, other),
                         KGCN_ERR_BAD_SHAPE, "kgcn_tfexample_gather: feature '%s' of record %lld is malformed", key,
                         (long long)r);
            KGCN_REQUIRE(!(other && local == 0), KGCN_ERR_UNSUPPORTED,
                         "kgcn_tfexample_gather: feature '%s' of record %lld holds another kind than requested (%d)",
                         key, (long long)r, kind);
            got = local;
        }
        if (counts) counts[r] = got;
        n += got;
    }
    *total = n;
    if (out != nullptr && n > capacity)
        return kgcn::fail(KGCN_ERR_WORKSPACE, "kgcn_tfexample_gather: '%s' needs room for %lld values, got %lld", key,
                          (long long)n, (long long)capacity);
    return KGCN_OK;
}
