// oracle_capi.cpp -- extern "C" surface over dcdf_oracle.hpp so tests (ctypes) and bench.py's
// cpu_baseline leg can drive the CPU restatement.  TEST INFRASTRUCTURE ONLY: the product library
// (dcdf_b200/csrc) never links or loads this.
#include <array>
#include <chrono>
#include <cstdio>
#include <thread>

#include "dcdf_oracle.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;

struct Obj {
  int kind = 0;  // 1 snapshot, 2 log (+snapshot), 3 chunk, 4 super
  std::unique_ptr<Snapshot> snap;
  std::unique_ptr<Log> log;
  std::shared_ptr<Chunk> chunk;
  std::shared_ptr<SuperNode> super;
  // flattened pre-order view of a superchunk tree
  struct Flat { const SuperNode* sn = nullptr; const Chunk* ch = nullptr; std::vector<int32_t> child; };
  std::vector<Flat> flat;
};

template <class Fn>
int32_t guard(Fn&& fn) {
  try {
    fn();
    return OK;
  } catch (const Error& e) {
    g_err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return BAD_ARG;
  }
}

Buffer3 make_buffer(int32_t enc, const void* base, const int64_t* shape, const int64_t* strides, int32_t fb, int32_t round) {
  Buffer3 b;
  b.enc = (Encoding)enc;
  b.base = base;
  for (int i = 0; i < 3; i++) { b.shape[i] = shape[i]; b.strides[i] = strides[i]; }
  b.fractional_bits = (usize)fb;
  b.round = round != 0;
  return b;
}

int32_t flatten(Obj* o, const SuperNode* sn) {
  int32_t me = (int32_t)o->flat.size();
  o->flat.emplace_back();
  o->flat[me].sn = sn;
  std::vector<int32_t> child(sn->refs.size(), -1);
  for (usize i = 0; i < sn->refs.size(); i++) {
    const SubRef& r = sn->refs[i];
    if (r.kind == 0) continue;
    if (r.chunk) {
      child[i] = (int32_t)o->flat.size();
      o->flat.emplace_back();
      o->flat.back().ch = r.chunk.get();
    } else {
      child[i] = flatten(o, r.super.get());
    }
  }
  o->flat[me].child = child;
  return me;
}

const BitMap& pick_bitmap(const Obj* o, int which) {
  if (o->kind == 1) { if (which == 0) return o->snap->nodemap; }
  if (o->kind == 2) return which == 0 ? o->log->nodemap : o->log->equal;
  fail(BAD_ARG, "no such bitmap");
}
const Dac& pick_dac(const Obj* o, int which) {
  if (o->kind == 1) return which == 0 ? o->snap->max : o->snap->min;
  if (o->kind == 2) return which == 0 ? o->log->max : o->log->min;
  fail(BAD_ARG, "no such dac");
}
}  // namespace

extern "C" {

struct dcdf_oracle_stats {
  uint64_t size;
  uint32_t elided, local, external, snapshots, logs;
};

const char* dcdf_oracle_last_error() { return g_err.c_str(); }
void dcdf_oracle_free(void* obj) { delete static_cast<Obj*>(obj); }

// ---- fixed.rs
int32_t dcdf_oracle_to_fixed_f32(float n, int32_t bits, int32_t round, int64_t* out) {
  return guard([&] { *out = to_fixed<float>(n, (usize)bits, round != 0); });
}
int32_t dcdf_oracle_to_fixed_f64(double n, int32_t bits, int32_t round, int64_t* out) {
  return guard([&] { *out = to_fixed<double>(n, (usize)bits, round != 0); });
}
float dcdf_oracle_from_fixed_f32(int64_t n, int32_t bits) { return from_fixed<float>(n, (usize)bits); }
double dcdf_oracle_from_fixed_f64(int64_t n, int32_t bits) { return from_fixed<double>(n, (usize)bits); }
void dcdf_oracle_from_fixed_array_f32(const int64_t* in, uint64_t n, int32_t bits, float* out) {
  for (uint64_t i = 0; i < n; i++) out[i] = from_fixed<float>(in[i], (usize)bits);
}
void dcdf_oracle_from_fixed_array_f64(const int64_t* in, uint64_t n, int32_t bits, double* out) {
  for (uint64_t i = 0; i < n; i++) out[i] = from_fixed<double>(in[i], (usize)bits);
}
int32_t dcdf_oracle_suggest_fraction(int32_t enc, const void* base, const int64_t* shape, const int64_t* strides,
                                     int32_t* kind, int32_t* bits) {
  return guard([&] {
    Buffer3 b = make_buffer(enc, base, shape, strides, 0, 0);
    Fraction f{false, 0};
    if (enc == ENC_F32) f = suggest_fraction(b.view<float>());
    else if (enc == ENC_F64) f = suggest_fraction(b.view<double>());
    else fail(BAD_ARG, "suggest_fraction needs a float array");
    *kind = f.round ? 1 : 0;
    *bits = (int32_t)f.bits;
  });
}
int32_t dcdf_oracle_min_max(int32_t enc, const void* base, const int64_t* shape, const int64_t* strides, int32_t fb,
                            int32_t round, int64_t* min_out, int64_t* max_out) {
  return guard([&] {
    Buffer3 b = make_buffer(enc, base, shape, strides, fb, round);
    auto mm = b.min_max();
    for (usize i = 0; i < mm.size(); i++) { min_out[i] = mm[i].first; max_out[i] = mm[i].second; }
  });
}

// ---- bitmap.rs / dac.rs primitives (known-answer tests)
int32_t dcdf_oracle_bitmap_from_bytes(const uint8_t* bytes, uint64_t n_bytes, uint64_t length, uint32_t* words_out,
                                      uint32_t* index_out, uint32_t* n_words, uint32_t* n_index) {
  return guard([&] {
    BitMapBuilder b;
    b.length = length;
    b.bytes.assign(bytes, bytes + n_bytes);
    BitMap m = b.finish();
    *n_words = (uint32_t)m.bitmap.size();
    *n_index = (uint32_t)m.index.size();
    for (usize i = 0; i < m.bitmap.size(); i++) words_out[i] = m.bitmap[i];
    for (usize i = 0; i < m.index.size(); i++) index_out[i] = m.index[i];
  });
}
// Push bits one at a time, then answer get/rank for every position 0..=length and serialize.
int32_t dcdf_oracle_bitmap_push_rank(const uint8_t* bits, uint64_t length, uint8_t* get_out, uint32_t* rank_out,
                                     uint8_t* ser_out, uint64_t ser_cap, uint64_t* ser_len) {
  return guard([&] {
    BitMapBuilder b;
    for (uint64_t i = 0; i < length; i++) b.push(bits[i] != 0);
    BitMap m = b.finish();
    for (uint64_t i = 0; i < length; i++) get_out[i] = m.get(i);
    for (uint64_t i = 0; i <= length; i++) rank_out[i] = (uint32_t)m.rank(i);
    Writer w;
    m.write_to(w);
    if (w.buf.size() != m.size()) fail(BAD_FORMAT, "bitmap len != size()");
    *ser_len = w.buf.size();
    if (w.buf.size() <= ser_cap) memcpy(ser_out, w.buf.data(), w.buf.size());
    Reader r{w.buf.data(), w.buf.size()};
    BitMap m2 = BitMap::read_from(r);
    for (uint64_t i = 0; i <= length; i++)
      if (m2.rank(i) != m.rank(i)) fail(BAD_FORMAT, "bitmap round trip mismatch");
  });
}
int32_t dcdf_oracle_dac_roundtrip(const int64_t* values, uint64_t n, int64_t* collect_out, uint8_t* ser_out,
                                  uint64_t ser_cap, uint64_t* ser_len, uint32_t* n_levels) {
  return guard([&] {
    Dac d = Dac::from(std::vector<i64>(values, values + n));
    for (uint64_t i = 0; i < n; i++) collect_out[i] = d.get(i);
    Writer w;
    d.write_to(w);
    if (w.buf.size() != d.size()) fail(BAD_FORMAT, "dac len != size()");
    *ser_len = w.buf.size();
    *n_levels = (uint32_t)d.levels.size();
    if (w.buf.size() <= ser_cap) memcpy(ser_out, w.buf.data(), w.buf.size());
    Reader r{w.buf.data(), w.buf.size()};
    Dac d2 = Dac::read_from(r);
    for (uint64_t i = 0; i < n; i++)
      if (d2.get(i) != values[i]) fail(BAD_FORMAT, "dac round trip mismatch");
  });
}
int64_t dcdf_oracle_dac_get_empty() { return Dac().get(0); }  // Appendix B #16

// ---- snapshot.rs / log.rs over plain i64 rasters (testing.rs:341-388 helpers)
int32_t dcdf_oracle_snapshot_build_i64(const int64_t* data, int64_t rows, int64_t cols, int32_t k, void** out) {
  return guard([&] {
    auto get = [&](usize r, usize c) { return data[r * cols + c]; };
    auto o = std::make_unique<Obj>();
    o->kind = 1;
    o->snap = std::make_unique<Snapshot>(Snapshot::build(get, (usize)rows, (usize)cols, k));
    *out = o.release();
  });
}
int32_t dcdf_oracle_log_build_i64(const int64_t* s, const int64_t* t, int64_t rows, int64_t cols, int32_t k, void** out) {
  return guard([&] {
    auto get_s = [&](usize r, usize c) { return s[r * cols + c]; };
    auto get_t = [&](usize r, usize c) { return t[r * cols + c]; };
    auto o = std::make_unique<Obj>();
    o->kind = 2;
    o->snap = std::make_unique<Snapshot>(Snapshot::build(get_s, (usize)rows, (usize)cols, k));
    o->log = std::make_unique<Log>(Log::build(get_s, get_t, (usize)rows, (usize)cols, k));
    *out = o.release();
  });
}
int32_t dcdf_oracle_bitmap_info(const void* obj, int32_t which, uint32_t* length, uint32_t* n_words, uint32_t* n_index) {
  return guard([&] {
    const BitMap& b = pick_bitmap(static_cast<const Obj*>(obj), which);
    *length = (uint32_t)b.length;
    *n_words = (uint32_t)b.bitmap.size();
    *n_index = (uint32_t)b.index.size();
  });
}
int32_t dcdf_oracle_bitmap_words(const void* obj, int32_t which, uint32_t* words, uint32_t* index) {
  return guard([&] {
    const BitMap& b = pick_bitmap(static_cast<const Obj*>(obj), which);
    for (usize i = 0; i < b.bitmap.size(); i++) words[i] = b.bitmap[i];
    for (usize i = 0; i < b.index.size(); i++) index[i] = b.index[i];
  });
}
int32_t dcdf_oracle_dac_len(const void* obj, int32_t which, uint64_t* n, uint32_t* n_levels) {
  return guard([&] {
    const Dac& d = pick_dac(static_cast<const Obj*>(obj), which);
    *n = d.len();
    *n_levels = (uint32_t)d.levels.size();
  });
}
int32_t dcdf_oracle_dac_collect(const void* obj, int32_t which, int64_t* out) {
  return guard([&] {
    const Dac& d = pick_dac(static_cast<const Obj*>(obj), which);
    auto v = d.collect();
    for (usize i = 0; i < v.size(); i++) out[i] = v[i];
  });
}
// Serialize a Snapshot (kind 1) or a Log (kind 2; which=1 serializes the log's snapshot instead).
int32_t dcdf_oracle_struct_serialize(const void* obj, int32_t which, uint8_t* out, uint64_t cap, uint64_t* len) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    Writer w;
    u64 size;
    if (o->kind == 1 || which == 1) { o->snap->write_to(w); size = o->snap->size(); }
    else { o->log->write_to(w); size = o->log->size(); }
    if (w.buf.size() != size) fail(BAD_FORMAT, "len != size()");
    *len = w.buf.size();
    if (w.buf.size() <= cap) memcpy(out, w.buf.data(), w.buf.size());
    Reader r{w.buf.data(), w.buf.size()};  // read back (snapshot.rs:873-890, log.rs:1620-1637)
    if (o->kind == 1 || which == 1) Snapshot::read_from(r); else Log::read_from(r);
    if (r.pos != w.buf.size()) fail(BAD_FORMAT, "trailing bytes");
  });
}
int32_t dcdf_oracle_struct_get(const void* obj, int64_t row, int64_t col, int64_t* out) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    *out = o->kind == 1 ? o->snap->get((usize)row, (usize)col) : o->log->get(*o->snap, (usize)row, (usize)col);
  });
}
int32_t dcdf_oracle_struct_window(const void* obj, int64_t top, int64_t bottom, int64_t left, int64_t right, int64_t* out) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    Rect b((usize)top, (usize)bottom, (usize)left, (usize)right);
    usize cols = b.cols();
    auto set = [&](usize r, usize c, i64 v) { out[r * cols + c] = v; };
    if (o->kind == 1) o->snap->fill_window(set, b); else o->log->fill_window(set, *o->snap, b);
  });
}
int32_t dcdf_oracle_struct_search(const void* obj, int64_t top, int64_t bottom, int64_t left, int64_t right,
                                  int64_t lower, int64_t upper, int64_t* rc_out, uint64_t cap, uint64_t* n) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    Rect b((usize)top, (usize)bottom, (usize)left, (usize)right);
    auto cells = o->kind == 1 ? o->snap->search_window(b, lower, upper) : o->log->search_window(*o->snap, b, lower, upper);
    *n = cells.size();
    for (usize i = 0; i < cells.size() && i < cap; i++) { rc_out[2 * i] = (i64)cells[i].first; rc_out[2 * i + 1] = (i64)cells[i].second; }
  });
}

// ---- chunk.rs
int32_t dcdf_oracle_chunk_build(int32_t enc, const void* base, const int64_t* shape, const int64_t* strides, int32_t k,
                                int32_t fractional_bits, int32_t round, void** out, dcdf_oracle_stats* stats) {
  return guard([&] {
    Buffer3 b = make_buffer(enc, base, shape, strides, fractional_bits, round);
    BuildStats st;
    auto o = std::make_unique<Obj>();
    o->kind = 3;
    o->chunk = std::make_shared<Chunk>(Chunk::build(b, (usize)shape[0], (usize)shape[1], (usize)shape[2], k, &st));
    if (stats) *stats = {st.size, (uint32_t)st.elided, (uint32_t)st.local, (uint32_t)st.external, (uint32_t)st.snapshots, (uint32_t)st.logs};
    *out = o.release();
  });
}
int32_t dcdf_oracle_chunk_open(const uint8_t* bytes, uint64_t len, void** out) {
  return guard([&] {
    Reader r{bytes, (usize)len};
    auto o = std::make_unique<Obj>();
    o->kind = 3;
    o->chunk = std::make_shared<Chunk>(Chunk::read_from(r));
    if (r.pos != len) fail(BAD_FORMAT, "trailing bytes after chunk");
    *out = o.release();
  });
}
static const Chunk* as_chunk(const void* obj, int32_t node) {
  const Obj* o = static_cast<const Obj*>(obj);
  if (o->kind == 3) return o->chunk.get();
  if (o->kind == 4 && node >= 0 && (usize)node < o->flat.size() && o->flat[node].ch) return o->flat[node].ch;
  fail(BAD_ARG, "not a chunk");
}
int32_t dcdf_oracle_chunk_serialize(const void* obj, int32_t node, uint8_t* out, uint64_t cap, uint64_t* len) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    Writer w;
    c->write_to(w);
    if (w.buf.size() != c->size()) fail(BAD_FORMAT, "chunk len != size()");
    *len = w.buf.size();
    if (w.buf.size() <= cap) memcpy(out, w.buf.data(), w.buf.size());
  });
}
int32_t dcdf_oracle_chunk_info(const void* obj, int32_t node, int64_t* shape, int32_t* enc, int32_t* fb, uint32_t* n_blocks) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    usize s[3];
    c->shape(s);
    for (int i = 0; i < 3; i++) shape[i] = (int64_t)s[i];
    *enc = c->encoding;
    *fb = (int32_t)c->fractional_bits;
    *n_blocks = (uint32_t)c->blocks.size();
  });
}
int32_t dcdf_oracle_chunk_block_instants(const void* obj, int32_t node, uint32_t* out) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    for (usize i = 0; i < c->blocks.size(); i++) out[i] = (uint32_t)(c->blocks[i].logs.size() + 1);
  });
}
static void check_cube(const Chunk* c, const Cube& b) {  // mmarray.rs:218-229 check_bounds
  usize s[3];
  c->shape(s);
  if (b.end > s[0] || b.bottom > s[1] || b.right > s[2]) fail(OUT_OF_BOUNDS, "window out of bounds");
}
int32_t dcdf_oracle_chunk_get_batch(const void* obj, int32_t node, uint64_t n, const int64_t* irc, int64_t* out) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    usize s[3];
    c->shape(s);
    for (uint64_t q = 0; q < n; q++) {
      usize i = (usize)irc[3 * q], r = (usize)irc[3 * q + 1], col = (usize)irc[3 * q + 2];
      if (i >= s[0] || r >= s[1] || col >= s[2]) fail(OUT_OF_BOUNDS, "cell out of bounds");
      out[q] = c->get(i, r, col);
    }
  });
}
int32_t dcdf_oracle_chunk_cell(const void* obj, int32_t node, int64_t start, int64_t end, int64_t row, int64_t col, int64_t* out) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    check_cube(c, Cube((usize)start, (usize)end, (usize)row, (usize)row + 1, (usize)col, (usize)col + 1));
    c->fill_cell((usize)start, (usize)end, (usize)row, (usize)col, [&](usize i, i64 v) { out[i] = v; });
  });
}
int32_t dcdf_oracle_chunk_window(const void* obj, int32_t node, const int64_t* cube, int64_t* out) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    Cube b((usize)cube[0], (usize)cube[1], (usize)cube[2], (usize)cube[3], (usize)cube[4], (usize)cube[5]);
    check_cube(c, b);
    usize rows = b.rows(), cols = b.cols();
    c->fill_window(b, [&](usize i, usize r, usize col, i64 v) { out[(i * rows + r) * cols + col] = v; });
  });
}
int32_t dcdf_oracle_chunk_search(const void* obj, int32_t node, const int64_t* cube, int64_t lower, int64_t upper,
                                 int64_t* out_irc, uint64_t cap, uint64_t* n) {
  return guard([&] {
    const Chunk* c = as_chunk(obj, node);
    Cube b((usize)cube[0], (usize)cube[1], (usize)cube[2], (usize)cube[3], (usize)cube[4], (usize)cube[5]);
    check_cube(c, b);
    auto res = c->search(b, lower, upper);
    *n = res.size();
    for (usize i = 0; i < res.size() && i < cap; i++)
      for (int j = 0; j < 3; j++) out_irc[3 * i + j] = (int64_t)res[i][j];
  });
}

// ---- superchunk.rs (compute part).  compute_bits != 0 also performs dataset.rs:842.
int32_t dcdf_oracle_superchunk_build(int32_t enc, const void* base, const int64_t* shape, const int64_t* strides,
                                     const uint32_t* levels, uint32_t n_levels, int32_t k, int32_t fractional_bits,
                                     int32_t round, int32_t compute_bits, void** out) {
  return guard([&] {
    Buffer3 b = make_buffer(enc, base, shape, strides, fractional_bits, round);
    if (compute_bits) b.compute_fractional_bits();
    auto o = std::make_unique<Obj>();
    o->kind = 4;
    o->super = SuperNode::build(b, (usize)shape[0], (usize)shape[1], (usize)shape[2], levels, n_levels, k);
    flatten(o.get(), o->super.get());
    *out = o.release();
  });
}
struct dcdf_oracle_node_info {
  int32_t kind;  // 0 superchunk, 1 chunk
  int32_t encoding, fractional_bits;
  uint32_t levels;
  int64_t shape[3];
  int64_t sidelen, chunks_sidelen, subsidelen;
  uint32_t n_refs;
  uint64_t bytes0, bytes1, bytes2;  // chunk bytes | max dac, min dac
  dcdf_oracle_stats stats;
};
int32_t dcdf_oracle_super_n_nodes(const void* obj, uint32_t* n) {
  return guard([&] { *n = (uint32_t) static_cast<const Obj*>(obj)->flat.size(); });
}
int32_t dcdf_oracle_super_node_info(const void* obj, uint32_t node, dcdf_oracle_node_info* info) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    if (node >= o->flat.size()) fail(BAD_ARG, "bad node index");
    const auto& f = o->flat[node];
    memset(info, 0, sizeof(*info));
    if (f.ch) {
      info->kind = 1;
      usize s[3];
      f.ch->shape(s);
      for (int i = 0; i < 3; i++) info->shape[i] = (int64_t)s[i];
      info->encoding = f.ch->encoding;
      info->fractional_bits = (int32_t)f.ch->fractional_bits;
      info->bytes0 = f.ch->size();
    } else {
      const SuperNode* s = f.sn;
      info->kind = 0;
      for (int i = 0; i < 3; i++) info->shape[i] = (int64_t)s->shape[i];
      info->encoding = s->encoding;
      info->fractional_bits = (int32_t)s->fractional_bits;
      info->levels = s->levels;
      info->sidelen = (int64_t)s->sidelen;
      info->chunks_sidelen = (int64_t)s->chunks_sidelen;
      info->subsidelen = (int64_t)s->subsidelen;
      info->n_refs = (uint32_t)s->refs.size();
      info->bytes1 = s->max.size();
      info->bytes2 = s->min.size();
      info->stats = {s->stats.size, (uint32_t)s->stats.elided, (uint32_t)s->stats.local, (uint32_t)s->stats.external,
                     (uint32_t)s->stats.snapshots, (uint32_t)s->stats.logs};
    }
  });
}
int32_t dcdf_oracle_super_node_refs(const void* obj, uint32_t node, int32_t* kinds, int32_t* child) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    if (node >= o->flat.size() || !o->flat[node].sn) fail(BAD_ARG, "bad node index");
    const auto& f = o->flat[node];
    for (usize i = 0; i < f.sn->refs.size(); i++) { kinds[i] = f.sn->refs[i].kind; child[i] = f.child[i]; }
  });
}
int32_t dcdf_oracle_super_node_bytes(const void* obj, uint32_t node, int32_t which, uint8_t* out, uint64_t cap, uint64_t* len) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    if (node >= o->flat.size()) fail(BAD_ARG, "bad node index");
    const auto& f = o->flat[node];
    Writer w;
    if (f.ch) f.ch->write_to(w);
    else if (which == 1) f.sn->max.write_to(w);
    else f.sn->min.write_to(w);
    *len = w.buf.size();
    if (w.buf.size() <= cap) memcpy(out, w.buf.data(), w.buf.size());
  });
}
int32_t dcdf_oracle_super_get_batch(const void* obj, uint64_t n, const int64_t* irc, int64_t* out_fixed, int32_t* out_bits) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    for (uint64_t q = 0; q < n; q++) {
      usize bits = 0;
      out_fixed[q] = o->super->get((usize)irc[3 * q], (usize)irc[3 * q + 1], (usize)irc[3 * q + 2], &bits);
      out_bits[q] = (int32_t)bits;
    }
  });
}
// Window at superchunk level, already converted to the native float type (f32 only for now) or raw fixed.
int32_t dcdf_oracle_super_window_f32(const void* obj, const int64_t* cube, float* out) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    Cube b((usize)cube[0], (usize)cube[1], (usize)cube[2], (usize)cube[3], (usize)cube[4], (usize)cube[5]);
    usize rows = b.rows(), cols = b.cols();
    o->super->fill_window(b, [&](usize i, usize r, usize c, i64 v, usize fb) { out[(i * rows + r) * cols + c] = from_fixed<float>(v, fb); });
  });
}

// Raw fixed-point window at superchunk level + the fractional bits of every cell's source.
int32_t dcdf_oracle_super_window_raw(const void* obj, const int64_t* cube, int64_t* out) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    Cube b((usize)cube[0], (usize)cube[1], (usize)cube[2], (usize)cube[3], (usize)cube[4], (usize)cube[5]);
    usize rows = b.rows(), cols = b.cols();
    o->super->fill_window(b, [&](usize i, usize r, usize c, i64 v, usize) { out[(i * rows + r) * cols + c] = v; });
  });
}
// Superchunk::fill_cell batched: q = n x (start, end, row, col), out_off[n + 1] element offsets.  `sectors` (may be
// null): SURVEY 8d's sector model -- sum over the queries of the distinct 32-byte sectors of serialized bytes touched.
int32_t dcdf_oracle_super_cell_batch(const void* obj, uint64_t n, const int64_t* q, const uint64_t* out_off, int64_t* out_fixed,
                                     uint64_t* sectors) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    SectorTracker tr;
    if (sectors) { *sectors = 0; o->super->assign_offsets(); sector_tracker() = &tr; }
    try {
      for (uint64_t i = 0; i < n; i++) {
        usize start = (usize)q[4 * i], end = (usize)q[4 * i + 1];
        if (start > end) std::swap(start, end);
        i64* dst = out_fixed ? out_fixed + out_off[i] : nullptr;
        o->super->fill_cell(start, end, (usize)q[4 * i + 2], (usize)q[4 * i + 3], [&](usize t, i64 v, usize) { if (dst) dst[t] = v; });
        if (sectors) { *sectors += tr.distinct(); tr.clear(); }
      }
    } catch (...) { sector_tracker() = nullptr; throw; }
    sector_tracker() = nullptr;
  });
}
// Superchunk::search batched over n windows with per-window [lower, upper]; counts[n]; out_irc (may be null) receives
// the triplets window after window (up to cap).  `sectors` as above.
int32_t dcdf_oracle_super_search_batch(const void* obj, uint64_t n, const int64_t* cubes, const int64_t* lower, const int64_t* upper,
                                       uint64_t* counts, int64_t* out_irc, uint64_t cap, uint64_t* n_found, uint64_t* sectors) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    SectorTracker tr;
    if (sectors) { *sectors = 0; o->super->assign_offsets(); sector_tracker() = &tr; }
    uint64_t total = 0;
    try {
      for (uint64_t i = 0; i < n; i++) {
        const int64_t* c = cubes + 6 * i;
        Cube b((usize)c[0], (usize)c[1], (usize)c[2], (usize)c[3], (usize)c[4], (usize)c[5]);
        std::vector<std::array<usize, 3>> hits;
        if (b.instants() && b.rows() && b.cols()) o->super->search(b, lower[i], upper[i], hits);
        if (counts) counts[i] = hits.size();
        for (auto& h : hits) {
          if (out_irc && total < cap) { out_irc[3 * total] = (int64_t)h[0]; out_irc[3 * total + 1] = (int64_t)h[1]; out_irc[3 * total + 2] = (int64_t)h[2]; }
          total++;
        }
        if (sectors) { *sectors += tr.distinct(); tr.clear(); }
      }
    } catch (...) { sector_tracker() = nullptr; throw; }
    sector_tracker() = nullptr;
    if (n_found) *n_found = total;
  });
}
// Sector model of a full-extent window read: every serialized byte of every structure the window needs is read once
// (S_in of SURVEY 8d); here simply the stored bytes of the superchunk.
int32_t dcdf_oracle_super_stored_bytes(const void* obj, uint64_t* nbytes) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    *nbytes = o->super->stats.size;
  });
}

// ---- storage side: what Superchunk::build + Resolver::save put into the store (testing.rs MemoryMapper)
struct SavedObj {
  MemoryStore store;
  CidB root;
  BuildStats stats;
};
int32_t dcdf_oracle_super_save(const void* obj, void** out) {
  return guard([&] {
    const Obj* o = static_cast<const Obj*>(obj);
    auto* sv = new SavedObj();
    SuperNode::Saved r = o->super->save(sv->store);
    Writer w;
    write_header(w, NODE_MMSTRUCT3);
    w.bytes(r.body);
    sv->root = cid_for_bytes(w.buf);
    // the superchunk node itself is listed last (as the caller's resolver.save would store it)
    sv->store.objects.emplace_back(sv->root, std::move(w.buf));
    sv->store.types.push_back(NODE_SUPERCHUNK);
    sv->stats = r.stats;
    *out = sv;
  });
}
int32_t dcdf_oracle_saved_free(void* p) { delete static_cast<SavedObj*>(p); return OK; }
int32_t dcdf_oracle_saved_count(const void* p, uint32_t* n) { *n = (uint32_t)static_cast<const SavedObj*>(p)->store.objects.size(); return OK; }
int32_t dcdf_oracle_saved_node(const void* p, uint32_t i, uint8_t* cid, int32_t* type, uint8_t* bytes, uint64_t cap, uint64_t* len) {
  return guard([&] {
    const SavedObj* sv = static_cast<const SavedObj*>(p);
    if (i >= sv->store.objects.size()) fail(BAD_ARG, "node index out of range");
    const auto& o = sv->store.objects[i];
    if (cid) memcpy(cid, o.first.data(), 36);
    if (type) *type = sv->store.types[i];
    if (len) *len = o.second.size();
    if (bytes && cap >= o.second.size()) memcpy(bytes, o.second.data(), o.second.size());
  });
}
int32_t dcdf_oracle_saved_stats(const void* p, uint64_t* size, uint32_t* elided, uint32_t* external, uint32_t* snapshots, uint32_t* logs) {
  const SavedObj* sv = static_cast<const SavedObj*>(p);
  *size = sv->stats.size; *elided = (uint32_t)sv->stats.elided; *external = (uint32_t)sv->stats.external;
  *snapshots = (uint32_t)sv->stats.snapshots; *logs = (uint32_t)sv->stats.logs;
  return OK;
}
int32_t dcdf_oracle_sha256(const uint8_t* p, uint64_t n, uint8_t* out32) {
  OSha256 h;
  h.update(p, (usize)n);
  auto d = h.finish();
  memcpy(out32, d.data(), 32);
  return OK;
}

// ---- CPU baseline helper: encode `n_units` independent [T,r,c] sub-arrays of one strided raster with
// `threads` host threads (the reference itself never spawns: superchunk.rs:123-188; threads=1 is faithful).
// Returns total serialized bytes and elapsed seconds.
int32_t dcdf_oracle_bench_superchunk(int32_t enc, const void* base, const int64_t* shape, const int64_t* strides,
                                     const uint32_t* levels, uint32_t n_levels, int32_t k, int32_t fractional_bits,
                                     int32_t round, int32_t threads, int32_t repeats, double* seconds, uint64_t* out_bytes) {
  return guard([&] {
    (void)threads;
    double best = 1e300;
    uint64_t bytes = 0;
    for (int rep = 0; rep < repeats; rep++) {
      auto t0 = std::chrono::steady_clock::now();
      Buffer3 b = make_buffer(enc, base, shape, strides, fractional_bits, round);
      b.compute_fractional_bits();
      auto sn = SuperNode::build(b, (usize)shape[0], (usize)shape[1], (usize)shape[2], levels, n_levels, k);
      auto t1 = std::chrono::steady_clock::now();
      best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
      bytes = sn->stats.size;
    }
    *seconds = best;
    *out_bytes = bytes;
  });
}

}  // extern "C"
