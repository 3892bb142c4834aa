// qb_formats.cu — a segment's files as they lie on disk -> HBM storages.
//
// The GPU copy is a cache of what the reference keeps in a segment directory:
//   matrix.dat            dense/immutable_dense_vectors.rs:25-27,100-113   b"data" + count x dim x size_of::<T>() row-major, no padding
//   quantized.data        quantized/quantized_storage.rs:63-69             headerless rows of `quantized_vector_size` bytes
//   quantized.meta.json   serde_json of MetadataInt8 (encoded_vectors_u8.rs:84-91), PQ Metadata (encoded_vectors_pq.rs:46-51) or
//                         BQ Metadata (encoded_vectors_binary.rs:112-125), each holding VectorParameters (encoded_vectors.rs:30-41)
// so that a maintainer hands the mmapped bytes over unchanged.  Host-side parsing only; uploads go through qb_storage_create_*.
#include <ctype.h>
#include <stdlib.h>

#include <map>
#include <memory>

#include "qb_internal.h"

namespace {

// ---------------------------------------------------------------- a small JSON reader (objects, arrays, numbers, strings, bool, null)
struct JVal {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    bool b = false;
    double num = 0.0;
    std::string text;                       // STR: the string; NUM: the literal as written (re-parsed with strtof for f32 fields)
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* key) const {
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p; const char* end; std::string err;
    void ws() { while (p < end && isspace((unsigned char)*p)) ++p; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    bool parse_string(std::string& out) {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) return fail("bad escape");
                switch (*p) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': if (end - p < 5) return fail("bad \\u escape"); out += '?'; p += 4; break;
                    default: out += *p;
                }
                ++p;
            } else out += *p++;
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool parse(JVal& v, int depth = 0) {
        if (depth > 32) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        if (*p == '{') {
            v.kind = JVal::OBJ; ++p; ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!parse_string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                v.obj.emplace_back(std::move(k), JVal());
                if (!parse(v.obj.back().second, depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (*p == '[') {
            v.kind = JVal::ARR; ++p; ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                v.arr.emplace_back();
                if (!parse(v.arr.back(), depth + 1)) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (*p == '"') { v.kind = JVal::STR; return parse_string(v.text); }
        if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JVal::BOOL; v.b = true; p += 4; return true; }
        if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JVal::BOOL; v.b = false; p += 5; return true; }
        if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JVal::NUL; p += 4; return true; }
        const char* s = p;
        while (p < end && (isdigit((unsigned char)*p) || *p == '-' || *p == '+' || *p == '.' || *p == 'e' || *p == 'E')) ++p;
        if (p == s) return fail("unexpected character");
        v.kind = JVal::NUM; v.text.assign(s, p - s); v.num = strtod(v.text.c_str(), nullptr);
        return true;
    }
};

// f32 fields: serde_json writes the shortest decimal that round-trips the f32; strtof of that text (or of a longer f64-style rendering
// of the same value) gives the f32 back exactly
bool as_f32(const JVal* v, float* out) { if (!v || v->kind != JVal::NUM) return false; *out = strtof(v->text.c_str(), nullptr); return true; }
bool as_u64(const JVal* v, uint64_t* out) { if (!v || v->kind != JVal::NUM || v->num < 0) return false; *out = (uint64_t)v->num; return true; }

struct VecParams { uint64_t dim = 0; int dt = -1; int invert = 0; };
bool vector_parameters(const JVal* v, VecParams* out) {   // encoded_vectors.rs:30-41
    if (!v || v->kind != JVal::OBJ) return false;
    if (!as_u64(v->get("dim"), &out->dim)) return false;
    const JVal* d = v->get("distance_type");
    if (!d || d->kind != JVal::STR) return false;
    if (d->text == "Cosine") out->dt = QB_QD_COSINE; else if (d->text == "Dot") out->dt = QB_QD_DOT; else if (d->text == "L1") out->dt = QB_QD_L1;
    else if (d->text == "L2") out->dt = QB_QD_L2; else return false;
    const JVal* i = v->get("invert");
    if (!i || i->kind != JVal::BOOL) return false;
    out->invert = i->b ? 1 : 0;
    return true;
}

}  // namespace

extern "C" qb_status qb_storage_load_dense_file(int32_t device, qb_dtype dt, qb_distance distance, uint32_t dim, const uint8_t* file_bytes, uint64_t n_bytes,
                                                qb_storage** out) {
    QB_CHECK(out, QB_ERR_INVALID, "load_dense_file: null out");
    *out = nullptr;
    QB_CHECK(file_bytes && n_bytes >= 4, QB_ERR_INVALID, "load_dense_file: %llu bytes is smaller than the header", (unsigned long long)n_bytes);
    QB_CHECK(!memcmp(file_bytes, "data", 4), QB_ERR_INVALID, "load_dense_file: header is not b\"data\" (dense_vector_storage.rs:31)");
    QB_CHECK(dim >= 1 && (int)dt >= 0 && (int)dt <= 2, QB_ERR_INVALID, "load_dense_file: bad dim / datatype");
    const uint64_t row = (uint64_t)dim * (dt == QB_DT_F32 ? 4 : (dt == QB_DT_F16 ? 2 : 1));
    // mmap-backed files are preallocated in whole pages: the vector count comes from the caller's bookkeeping in the reference; here every
    // complete row after the header is taken
    const uint64_t count = (n_bytes - 4) / row;
    return qb_storage_create_dense(device, dt, distance, dim, count, file_bytes + 4, row, out);
}

extern "C" qb_status qb_storage_load_quantized(int32_t device, qb_distance metric, const char* meta_json, uint64_t json_len, const uint8_t* data, uint64_t n_bytes,
                                               uint64_t count, qb_storage** out) {
    QB_CHECK(out, QB_ERR_INVALID, "load_quantized: null out");
    *out = nullptr;
    QB_CHECK(meta_json && json_len, QB_ERR_INVALID, "load_quantized: null metadata");
    QB_CHECK(data || n_bytes == 0, QB_ERR_INVALID, "load_quantized: null data");
    JParser jp{meta_json, meta_json + json_len, {}};
    JVal root;
    QB_CHECK(jp.parse(root) && root.kind == JVal::OBJ, QB_ERR_INVALID, "load_quantized: quantized.meta.json: %s", jp.err.empty() ? "not an object" : jp.err.c_str());
    VecParams vp;
    QB_CHECK(vector_parameters(root.get("vector_parameters"), &vp), QB_ERR_INVALID, "load_quantized: missing or malformed vector_parameters");
    QB_CHECK(vp.dim >= 1 && vp.dim <= 65536, QB_ERR_INVALID, "load_quantized: dim %llu", (unsigned long long)vp.dim);
    const uint32_t dim = (uint32_t)vp.dim;
    if (root.get("actual_dim")) {                                   // MetadataInt8, encoded_vectors_u8.rs:84-91
        uint64_t ad = 0; float alpha, offset, mult;
        QB_CHECK(as_u64(root.get("actual_dim"), &ad) && as_f32(root.get("alpha"), &alpha) && as_f32(root.get("offset"), &offset) && as_f32(root.get("multiplier"), &mult),
                 QB_ERR_INVALID, "load_quantized: malformed scalar-quantization metadata");
        const uint64_t stride = ad + 4;
        if (!count) count = n_bytes / stride;
        QB_CHECK(count * stride <= n_bytes, QB_ERR_INVALID, "load_quantized: %llu rows of %llu bytes exceed quantized.data (%llu bytes)", (unsigned long long)count,
                 (unsigned long long)stride, (unsigned long long)n_bytes);
        return qb_storage_create_sq8(device, dim, count, data, (uint32_t)stride, alpha, offset, mult, (qb_qdistance)vp.dt, vp.invert, metric, out);
    }
    if (const JVal* cents = root.get("centroids")) {               // PQ Metadata, encoded_vectors_pq.rs:46-51
        const JVal* div = root.get("vector_division");
        QB_CHECK(cents->kind == JVal::ARR && !cents->arr.empty() && div && div->kind == JVal::ARR && !div->arr.empty(), QB_ERR_INVALID,
                 "load_quantized: malformed product-quantization metadata");
        const uint32_t nc = (uint32_t)cents->arr.size(), m = (uint32_t)div->arr.size();
        std::vector<float> c((size_t)nc * dim);
        for (uint32_t i = 0; i < nc; ++i) {
            const JVal& row = cents->arr[i];
            QB_CHECK(row.kind == JVal::ARR && row.arr.size() == dim, QB_ERR_INVALID, "load_quantized: centroid %u has %zu values, dim is %u", i, row.arr.size(), dim);
            for (uint32_t k = 0; k < dim; ++k) QB_CHECK(as_f32(&row.arr[k], &c[(size_t)i * dim + k]), QB_ERR_INVALID, "load_quantized: centroid value is not a number");
        }
        std::vector<uint32_t> d(2 * (size_t)m);
        for (uint32_t j = 0; j < m; ++j) {                          // Range<usize> serialises as {"start":..,"end":..}
            uint64_t s = 0, e = 0;
            QB_CHECK(as_u64(div->arr[j].get("start"), &s) && as_u64(div->arr[j].get("end"), &e), QB_ERR_INVALID, "load_quantized: malformed vector_division[%u]", j);
            d[2 * j] = (uint32_t)s; d[2 * j + 1] = (uint32_t)e;
        }
        if (!count) count = n_bytes / m;
        QB_CHECK(count * m <= n_bytes, QB_ERR_INVALID, "load_quantized: %llu rows of %u bytes exceed quantized.data", (unsigned long long)count, m);
        return qb_storage_create_pq(device, dim, m, d.data(), c.data(), nc, data, count, (qb_qdistance)vp.dt, vp.invert, metric, out);
    }
    // BQ Metadata, encoded_vectors_binary.rs:112-125 (encoding / query_encoding / vector_stats are omitted when default)
    int enc = QB_BQ_ONE_BIT, qenc = QB_BQQ_SAME_AS_STORAGE;
    if (const JVal* e = root.get("encoding")) {
        QB_CHECK(e->kind == JVal::STR, QB_ERR_INVALID, "load_quantized: malformed encoding");
        if (e->text == "OneBit") enc = QB_BQ_ONE_BIT; else if (e->text == "TwoBits") enc = QB_BQ_TWO_BITS; else if (e->text == "OneAndHalfBits") enc = QB_BQ_ONE_AND_HALF_BITS;
        else { qb_set_error("load_quantized: unknown encoding '%s'", e->text.c_str()); return QB_ERR_INVALID; }
    }
    if (const JVal* e = root.get("query_encoding")) {
        QB_CHECK(e->kind == JVal::STR, QB_ERR_INVALID, "load_quantized: malformed query_encoding");
        if (e->text == "SameAsStorage") qenc = QB_BQQ_SAME_AS_STORAGE; else if (e->text == "Scalar4bits") qenc = QB_BQQ_SCALAR4; else if (e->text == "Scalar8bits") qenc = QB_BQQ_SCALAR8;
        else { qb_set_error("load_quantized: unknown query_encoding '%s'", e->text.c_str()); return QB_ERR_INVALID; }
    }
    std::vector<float> ms;
    if (const JVal* st = root.get("vector_stats")) {
        if (st->kind == JVal::OBJ) {
            const JVal* el = st->get("elements_stats");
            QB_CHECK(el && el->kind == JVal::ARR && el->arr.size() == dim, QB_ERR_INVALID, "load_quantized: vector_stats does not hold %u elements", dim);
            ms.resize((size_t)dim * 2);
            for (uint32_t k = 0; k < dim; ++k)
                QB_CHECK(as_f32(el->arr[k].get("mean"), &ms[2 * k]) && as_f32(el->arr[k].get("stddev"), &ms[2 * k + 1]), QB_ERR_INVALID, "load_quantized: malformed vector_stats[%u]", k);
        }
    }
    QB_CHECK(enc == QB_BQ_ONE_BIT || !ms.empty(), QB_ERR_INVALID, "load_quantized: two-bit / one-and-a-half-bit encodings need vector_stats");
    uint32_t rb = qb_bq_row_bytes(dim, (qb_bq_encoding)enc);
    {   // multivector storages hold u8-word rows: ceil(bits / 8) bytes (quantized_vectors.rs:270-282); recognised when the byte count says so
        uint64_t ext = dim;
        if (enc == QB_BQ_TWO_BITS) ext = (uint64_t)dim * 2; else if (enc == QB_BQ_ONE_AND_HALF_BITS) ext = ((uint64_t)dim * 3 + 1) / 2;
        const uint32_t rb8 = (uint32_t)((ext + 7) / 8);
        if (count && rb8 != rb && n_bytes < count * (uint64_t)rb && n_bytes >= count * (uint64_t)rb8) rb = rb8;
    }
    if (!count) count = n_bytes / rb;
    QB_CHECK(count * rb <= n_bytes, QB_ERR_INVALID, "load_quantized: %llu rows of %u bytes exceed quantized.data", (unsigned long long)count, rb);
    return qb_storage_create_bq(device, dim, (qb_bq_encoding)enc, (qb_bq_query_encoding)qenc, data, rb, count, (qb_qdistance)vp.dt, vp.invert, ms.empty() ? nullptr : ms.data(),
                                metric, out);
}
