// MessagePack reader / writer for the two SEND-SLAM wire messages next to the ORB hot path (include/orbx_wire.h).
// Also home of the binary-PNM header reader (orbx_pnm_header, declared in include/orbx.h): everything in this file parses untrusted
// bytes and is host code only, so that it can be built and fuzzed on its own (tests/test_wire_fuzz_cpu.py: ASan + UBSan).  The reader walks the payload in place (no object tree, no allocation); the subset of MessagePack it
// understands is the whole format, because unknown keys of any type have to be skipped the way msgpack-c's unpack accepts them.
#include <cstdint>
#include <cstring>

#include "../../include/orbx_wire.h"

namespace {

struct Reader {
    const uint8_t *p;
    size_t n, pos = 0;
    bool ok = true;

    bool need(size_t k) { if (!ok || n - pos < k) { ok = false; return false; } return true; }
    uint64_t be(size_t k) {                       // big-endian unsigned of k bytes
        if (!need(k)) return 0;
        uint64_t v = 0;
        for (size_t i = 0; i < k; i++) v = (v << 8) | p[pos + i];
        pos += k;
        return v;
    }
    int tag() { return need(1) ? p[pos++] : -1; }
    void skip_bytes(uint64_t k) { if (ok && k <= n - pos) pos += (size_t)k; else ok = false; }
};

enum Kind { K_NIL, K_BOOL, K_UINT, K_INT, K_F32, K_F64, K_STR, K_BIN, K_ARRAY, K_MAP, K_EXT, K_BAD };

struct Value {
    Kind kind = K_BAD;
    uint64_t u = 0;           // K_UINT value, K_BOOL, container / string length
    int64_t i = 0;            // K_INT (negative)
    double f = 0;             // K_F32 / K_F64
    const uint8_t *data = nullptr;   // K_STR / K_BIN / K_EXT body
};

// Reads one value header (and, for scalars / strings / bins, the whole value); containers are left for the caller to walk.
Value read_value(Reader &r) {
    Value v;
    const int t = r.tag();
    if (t < 0) return v;
    auto body = [&](Kind k, uint64_t len) { v.kind = k; v.u = len; v.data = r.p + r.pos; r.skip_bytes(len); };
    if (t <= 0x7f) { v.kind = K_UINT; v.u = (uint64_t)t; }
    else if (t <= 0x8f) { v.kind = K_MAP; v.u = (uint64_t)(t & 0x0f); }
    else if (t <= 0x9f) { v.kind = K_ARRAY; v.u = (uint64_t)(t & 0x0f); }
    else if (t <= 0xbf) body(K_STR, (uint64_t)(t & 0x1f));
    else if (t >= 0xe0) { v.kind = K_INT; v.i = (int8_t)t; }
    else switch (t) {
        case 0xc0: v.kind = K_NIL; break;
        case 0xc2: case 0xc3: v.kind = K_BOOL; v.u = (uint64_t)(t & 1); break;
        case 0xc4: body(K_BIN, r.be(1)); break;
        case 0xc5: body(K_BIN, r.be(2)); break;
        case 0xc6: body(K_BIN, r.be(4)); break;
        case 0xc7: { const uint64_t len = r.be(1); r.skip_bytes(1); body(K_EXT, len); break; }
        case 0xc8: { const uint64_t len = r.be(2); r.skip_bytes(1); body(K_EXT, len); break; }
        case 0xc9: { const uint64_t len = r.be(4); r.skip_bytes(1); body(K_EXT, len); break; }
        case 0xca: { const uint32_t b = (uint32_t)r.be(4); float f; std::memcpy(&f, &b, 4); v.kind = K_F32; v.f = f; break; }
        case 0xcb: { const uint64_t b = r.be(8); std::memcpy(&v.f, &b, 8); v.kind = K_F64; break; }
        case 0xcc: v.kind = K_UINT; v.u = r.be(1); break;
        case 0xcd: v.kind = K_UINT; v.u = r.be(2); break;
        case 0xce: v.kind = K_UINT; v.u = r.be(4); break;
        case 0xcf: v.kind = K_UINT; v.u = r.be(8); break;
        // signed families: a non-negative value is a "positive integer" for msgpack-c as well
        case 0xd0: { const int64_t s = (int8_t)r.be(1); if (s < 0) { v.kind = K_INT; v.i = s; } else { v.kind = K_UINT; v.u = (uint64_t)s; } break; }
        case 0xd1: { const int64_t s = (int16_t)r.be(2); if (s < 0) { v.kind = K_INT; v.i = s; } else { v.kind = K_UINT; v.u = (uint64_t)s; } break; }
        case 0xd2: { const int64_t s = (int32_t)r.be(4); if (s < 0) { v.kind = K_INT; v.i = s; } else { v.kind = K_UINT; v.u = (uint64_t)s; } break; }
        case 0xd3: { const int64_t s = (int64_t)r.be(8); if (s < 0) { v.kind = K_INT; v.i = s; } else { v.kind = K_UINT; v.u = (uint64_t)s; } break; }
        case 0xd4: r.skip_bytes(1); body(K_EXT, 1); break;
        case 0xd5: r.skip_bytes(1); body(K_EXT, 2); break;
        case 0xd6: r.skip_bytes(1); body(K_EXT, 4); break;
        case 0xd7: r.skip_bytes(1); body(K_EXT, 8); break;
        case 0xd8: r.skip_bytes(1); body(K_EXT, 16); break;
        case 0xd9: body(K_STR, r.be(1)); break;
        case 0xda: body(K_STR, r.be(2)); break;
        case 0xdb: body(K_STR, r.be(4)); break;
        case 0xdc: v.kind = K_ARRAY; v.u = r.be(2); break;
        case 0xdd: v.kind = K_ARRAY; v.u = r.be(4); break;
        case 0xde: v.kind = K_MAP; v.u = r.be(2); break;
        case 0xdf: v.kind = K_MAP; v.u = r.be(4); break;
        default: r.ok = false; break;             // 0xc1: never used
    }
    if (!r.ok) v.kind = K_BAD;
    return v;
}

// Skips the rest of a value whose header has been read: iterative, a counter of values still owed instead of recursion.
void skip_rest(Reader &r, const Value &head) {
    uint64_t owed = head.kind == K_MAP ? 2 * head.u : head.kind == K_ARRAY ? head.u : 0;
    while (owed && r.ok) {
        const Value v = read_value(r);
        owed--;
        if (v.kind == K_MAP) owed += 2 * v.u;
        else if (v.kind == K_ARRAY) owed += v.u;
        if (owed > r.n) r.ok = false;             // more values than bytes: malformed
    }
}

bool key_is(const Value &k, const char *name) {
    const size_t len = std::strlen(name);
    return k.u == len && std::memcmp(k.data, name, len) == 0;
}

bool to_double(const Value &v, double *out) {     // msgpack-c convert<double>: floats and integers
    switch (v.kind) {
        case K_F32: case K_F64: *out = v.f; return true;
        case K_UINT: *out = (double)v.u; return true;
        case K_INT: *out = (double)v.i; return true;
        default: return false;
    }
}

bool to_int(const Value &v, int *out) {           // msgpack-c convert<int>: integers in range
    if (v.kind == K_UINT && v.u <= 2147483647ull) { *out = (int)v.u; return true; }
    if (v.kind == K_INT && v.i >= -2147483648ll) { *out = (int)v.i; return true; }
    return false;
}

struct Writer {
    uint8_t *p;
    size_t cap, pos = 0;
    bool ok = true;
    void raw(const void *src, size_t k) { if (ok && k <= cap - pos) { if (k) std::memcpy(p + pos, src, k); pos += k; } else ok = false; }
    void byte(uint8_t b) { raw(&b, 1); }
    void be(uint64_t v, int k) { for (int i = k - 1; i >= 0; i--) byte((uint8_t)(v >> (8 * i))); }
    void str(const char *s) { const size_t len = std::strlen(s); byte((uint8_t)(0xa0 | len)); raw(s, len); }   // names here are < 32 bytes
    void integer(int64_t v) {
        if (v >= 0 && v <= 0x7f) byte((uint8_t)v);
        else if (v < 0 && v >= -32) byte((uint8_t)v);
        else if (v >= 0 && v <= 0xffff) { byte(0xcd); be((uint64_t)v, 2); }
        else if (v >= 0) { byte(0xce); be((uint64_t)v, 4); }
        else { byte(0xd2); be((uint64_t)(uint32_t)(int32_t)v, 4); }
    }
    void f64(double d) { uint64_t b; std::memcpy(&b, &d, 8); byte(0xcb); be(b, 8); }
    void bin(const void *src, size_t len) { byte(0xc6); be(len, 4); raw(src, len); }
};

}  // namespace

namespace {
// Byte cursor with OpenCV's PxM header grammar: a number is preceded by any run of white space / '#' comments and followed by
// exactly one consumed byte (so "255\n" leaves the cursor on the first sample).
struct PnmCursor {
    const uint8_t *d; size_t n, pos = 0; bool eos = false;
    int get() { if (pos >= n) { eos = true; return -1; } return d[pos++]; }
    static bool space(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
    static bool digit(int c) { return c >= '0' && c <= '9'; }
    bool number(long long &out) {
        int c = get();
        while (!eos && !digit(c)) {
            if (c == '#') { do c = get(); while (!eos && c != '\n' && c != '\r'); c = get(); }
            else if (space(c)) { do c = get(); while (!eos && space(c)); }
            else return false;
        }
        if (eos) return false;
        long long v = 0;
        while (true) {
            v = v * 10 + (c - '0');
            if (v > INT32_MAX) return false;
            c = get();
            if (eos) return false;            // the reader needs the terminating byte
            if (!digit(c)) break;
        }
        out = v;
        return true;
    }
};
}  // namespace

extern "C" {

int orbx_pnm_header(const uint8_t *data, size_t nbytes, int *width, int *height, int *channels, size_t *payload_offset) {
    if ((!data && nbytes) || !width || !height || !channels || !payload_offset) return ORBX_E_INVALID;
    if (nbytes < 2 || data[0] != 'P' || data[1] < '1' || data[1] > '6') return ORBX_E_EMPTY;   // imdecode: empty Mat
    PnmCursor cur{data, nbytes, 2};
    const int type = data[1] - '0';
    long long w = 0, ht = 0, maxval = 1;
    if (!cur.number(w) || !cur.number(ht)) return ORBX_E_EMPTY;
    if (type != 1 && type != 4 && !cur.number(maxval)) return ORBX_E_EMPTY;
    if (w <= 0 || ht <= 0 || maxval <= 0 || maxval > 65535) return ORBX_E_EMPTY;
    if ((type != 5 && type != 6) || maxval > 255) return ORBX_E_INVALID;                          // decodable, but not CV_8U binary
    const int ch = type == 6 ? 3 : 1;
    if ((unsigned long long)w * (unsigned long long)ht * ch > nbytes - cur.pos) return ORBX_E_EMPTY;   // truncated payload
    *width = (int)w; *height = (int)ht; *channels = ch; *payload_offset = cur.pos;
    return ORBX_OK;
}

int orbx_wire_parse_frame(const uint8_t *payload, size_t nbytes, orbx_wire_frame *out) {
    if (!payload || !out) return ORBX_E_INVALID;
    std::memset(out, 0, sizeof(*out));
    Reader r{payload, nbytes};
    const Value root = read_value(r);
    if (!r.ok || root.kind != K_MAP) return ORBX_E_INVALID;
    for (uint64_t e = 0; e < root.u; e++) {
        const Value key = read_value(r);
        if (!r.ok || (key.kind != K_STR && key.kind != K_BIN)) return ORBX_E_INVALID;   // key.convert(std::string) throws otherwise
        const Value val = read_value(r);
        if (!r.ok) return ORBX_E_INVALID;
        if (key_is(key, "type")) {
            if (val.kind != K_STR && val.kind != K_BIN) return ORBX_E_INVALID;
            out->type = val.data; out->type_len = (size_t)val.u;
        } else if (key_is(key, "timestamp")) {
            if (!to_double(val, &out->timestamp)) return ORBX_E_INVALID;
            out->has_timestamp = 1;
        } else if (key_is(key, "image") || key_is(key, "frame")) {
            if (val.kind != K_BIN) return ORBX_E_INVALID;          // "Image data must be encoded as MessagePack bin"
            out->image = val.data; out->image_bytes = (size_t)val.u;
        } else if (key_is(key, "camera_id")) {
            if (!to_int(val, &out->camera_id)) return ORBX_E_INVALID;
            out->has_camera_id = 1;
        } else {
            skip_rest(r, val);                                      // calibration sections and unknown fields
            if (!r.ok) return ORBX_E_INVALID;
        }
    }
    return out->type && out->type_len ? ORBX_OK : ORBX_E_INVALID;    // ParseMessage: return !packet.type.empty()
}

size_t orbx_wire_features_bound(int n, int framed) {
    const size_t nn = n > 0 ? (size_t)n : 0;
    return (framed ? 4 : 0) + 160 + nn * (sizeof(orbx_keypoint) + ORBX_DESC_BYTES);
}

int orbx_wire_pack_features(double timestamp, int camera_id, int width, int height, int mono_index, const orbx_keypoint *kp,
                            const uint8_t *desc, int n, int framed, uint8_t *out, size_t out_cap, size_t *written) {
    if (!out || !written || n < 0 || (n > 0 && (!kp || !desc))) return ORBX_E_INVALID;
    Writer w{out, out_cap};
    if (framed) w.be(0, 4);
    w.byte(0x89);                                                    // fixmap, 9 entries
    w.str("type"); w.str("features");
    w.str("camera_id"); w.integer(camera_id);
    w.str("timestamp"); w.f64(timestamp);
    w.str("width"); w.integer(width);
    w.str("height"); w.integer(height);
    w.str("mono_index"); w.integer(mono_index);
    w.str("n"); w.integer(n);
    w.str("keypoints"); w.bin(kp, (size_t)n * sizeof(orbx_keypoint));
    w.str("descriptors"); w.bin(desc, (size_t)n * ORBX_DESC_BYTES);
    if (!w.ok) return ORBX_E_CAPACITY;
    if (framed) {
        const uint64_t len = w.pos - 4;
        if (len > 0xffffffffull) return ORBX_E_CAPACITY;
        for (int i = 0; i < 4; i++) out[i] = (uint8_t)(len >> (8 * (3 - i)));
    }
    *written = w.pos;
    return ORBX_OK;
}

int orbx_wire_parse_features(const uint8_t *payload, size_t nbytes, orbx_wire_features *out) {
    if (!payload || !out) return ORBX_E_INVALID;
    std::memset(out, 0, sizeof(*out));
    Reader r{payload, nbytes};
    const Value root = read_value(r);
    if (!r.ok || root.kind != K_MAP) return ORBX_E_INVALID;
    bool is_features = false, have_n = false;
    size_t kp_bytes = 0, desc_bytes = 0;
    out->mono_index = -1;
    for (uint64_t e = 0; e < root.u; e++) {
        const Value key = read_value(r);
        if (!r.ok || (key.kind != K_STR && key.kind != K_BIN)) return ORBX_E_INVALID;
        const Value val = read_value(r);
        if (!r.ok) return ORBX_E_INVALID;
        if (key_is(key, "type")) is_features = (val.kind == K_STR || val.kind == K_BIN) && key_is(val, "features");
        else if (key_is(key, "timestamp")) { if (!to_double(val, &out->timestamp)) return ORBX_E_INVALID; }
        else if (key_is(key, "camera_id")) { if (!to_int(val, &out->camera_id)) return ORBX_E_INVALID; }
        else if (key_is(key, "width")) { if (!to_int(val, &out->width)) return ORBX_E_INVALID; }
        else if (key_is(key, "height")) { if (!to_int(val, &out->height)) return ORBX_E_INVALID; }
        else if (key_is(key, "mono_index")) { if (!to_int(val, &out->mono_index)) return ORBX_E_INVALID; }
        else if (key_is(key, "n")) { if (!to_int(val, &out->n)) return ORBX_E_INVALID; have_n = true; }
        else if (key_is(key, "keypoints")) {
            if (val.kind != K_BIN) return ORBX_E_INVALID;
            out->keypoints = val.data; kp_bytes = (size_t)val.u;
        } else if (key_is(key, "descriptors")) {
            if (val.kind != K_BIN) return ORBX_E_INVALID;
            out->descriptors = val.data; desc_bytes = (size_t)val.u;
        } else {
            skip_rest(r, val);
            if (!r.ok) return ORBX_E_INVALID;
        }
    }
    if (!is_features || !have_n || out->n < 0 || !out->keypoints || !out->descriptors) return ORBX_E_INVALID;
    if (kp_bytes != (size_t)out->n * sizeof(orbx_keypoint) || desc_bytes != (size_t)out->n * ORBX_DESC_BYTES) return ORBX_E_INVALID;
    return ORBX_OK;
}

int orbx_wire_copy_keypoints(const orbx_wire_features *f, orbx_keypoint *dst, int cap) {
    if (!f || f->n < 0 || (f->n > 0 && (!f->keypoints || !dst))) return ORBX_E_INVALID;
    if (cap < f->n) return ORBX_E_CAPACITY;
    if (f->n) std::memcpy(dst, f->keypoints, (size_t)f->n * sizeof(orbx_keypoint));
    return f->n;
}

}  // extern "C"
