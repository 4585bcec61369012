// api_store.cu -- the storage side of the boundary: content addressing + node assembly of built
// superchunks, and opening superchunks from STORED bytes so that the batched decoders serve data that was not built in
// this process.
//
//   Resolver::save                resolver.rs:126-138  (magic 0xDCE0, version 1, node type; then Node::save_to)
//   MMStruct3 node wrapper        mmstruct.rs:199-245  (NODE_SUBCHUNK + Chunk::write_to | NODE_SUPERCHUNK + save_to)
//   Superchunk::save_to / load_from   superchunk.rs:678-768, Reference :826-869
//   tail of Superchunk::build     superchunk.rs:199-270 (CID de-duplication of External references, Links, sizes)
//   Links::save_to / load_from    links.rs:41-76
//   CIDv1 of a SHA2-256 multihash testing.rs:172-183 (the reference's own in-memory store)
//
// Chunk nodes are hashed where they are (HBM, k_sha256: one thread per chunk); the Links and superchunk nodes are a
// few KB and are assembled and hashed on the host.  No CPU fallback for the chunk hashes.
#include <algorithm>
#include <array>
#include <functional>
#include <map>
#include <memory>
#include <mutex>

#include "host.hpp"
#include "sha256.cuh"
#include "tree_geo.hpp"

using namespace dcdf;

namespace dcdf {
void build_super_meta(dcdf_ctx* ctx, const dcdf_superchunk* sc);  // api_decode.cu: device directory + validation
}

namespace {

enum : uint8_t { NODE_LINKS = 1, NODE_MMSTRUCT3 = 2, NODE_SUBCHUNK = 4, NODE_SUPERCHUNK = 5 };  // node.rs:9-15

template <typename Fn>
int32_t guarded(dcdf_ctx* ctx, Fn&& fn) {
  if (!ctx) return DCDF_ERR_BAD_ARG;
  try {
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) throw CudaFail{std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
    fn();
    return DCDF_OK;
  } catch (const ApiFail& f) {
    ctx->last_error = f.msg;
    return f.code;
  } catch (const CudaFail& f) {
    ctx->last_error = f.msg;
    cudaGetLastError();
    return DCDF_ERR_CUDA;
  } catch (const std::exception& e) {
    ctx->last_error = std::string("unexpected exception: ") + e.what();
    return DCDF_ERR_BAD_ARG;
  } catch (...) {
    ctx->last_error = "unexpected exception";
    return DCDF_ERR_BAD_ARG;
  }
}

typedef std::array<uint8_t, DCDF_CID_BYTES> CidBytes;

// Cid::new_v1(0x12, Multihash::wrap(0x12 /* sha2-256 */, digest))  testing.rs:172-177: version, codec, hash code, length
CidBytes cid_of_digest(const uint8_t* digest) {
  CidBytes c;
  c[0] = 0x01; c[1] = 0x12; c[2] = 0x12; c[3] = 0x20;
  memcpy(c.data() + 4, digest, 32);
  return c;
}
CidBytes cid_of_bytes(const std::vector<uint8_t>& bytes) {
  Sha256 s;
  s.init();
  s.update(bytes.data(), bytes.size());
  uint8_t d[32];
  s.finish(d);
  return cid_of_digest(d);
}

void put_u32(std::vector<uint8_t>& v, uint32_t w) {  // extio.rs:228-233
  v.push_back((uint8_t)(w >> 24)); v.push_back((uint8_t)(w >> 16)); v.push_back((uint8_t)(w >> 8)); v.push_back((uint8_t)w);
}
void put_header(std::vector<uint8_t>& v, uint8_t node_type) {  // resolver.rs:130-132
  v.push_back(0xDC); v.push_back(0xE0);  // MAGIC_NUMBER = 0xDCDF + 1
  put_u32(v, 1);                         // FORMAT_VERSION
  v.push_back(node_type);
}

struct SavedNode {
  CidBytes cid;
  int32_t type;                 // NODE_SUBCHUNK / NODE_LINKS / NODE_SUPERCHUNK
  std::vector<uint8_t> bytes;   // whole stored object; for chunk nodes only the 8 bytes in front of the Chunk bytes
  uint64_t dev_off = 0, dev_len = 0;  // chunk nodes: Chunk bytes inside the superchunk's device blob
};

}  // namespace

struct dcdf_saved {
  int device = 0;
  const dcdf_superchunk* sc = nullptr;
  std::vector<SavedNode> nodes;  // distinct stored objects in first-save order; the last one is the superchunk node itself
  dcdf_build_stats stats;
};

namespace {

// SHA2-256 of every stored chunk node of the handle (all slices at once), cached in the handle.
std::mutex g_digest_mutex;
void ensure_digests(dcdf_ctx* ctx, const dcdf_superchunk* scc) {
  dcdf_superchunk* sc = const_cast<dcdf_superchunk*>(scc);
  std::lock_guard<std::mutex> lock(g_digest_mutex);
  if (!sc->unit_digests.empty()) return;
  const size_t n = sc->units.size();
  std::vector<HashJob> jobs;
  std::vector<uint32_t> job_unit;
  for (size_t u = 0; u < n; u++) {
    if (!sc->stored[u]) continue;
    HashJob j;
    memset(&j, 0, sizeof j);
    j.off = sc->chunk_off[u];
    j.len = sc->results[u].bytes;
    const uint8_t prefix[8] = {0xDC, 0xE0, 0, 0, 0, 1, NODE_MMSTRUCT3, NODE_SUBCHUNK};
    memcpy(j.prefix, prefix, 8);
    j.n_prefix = 8;
    jobs.push_back(j);
    job_unit.push_back((uint32_t)u);
  }
  sc->unit_digests.assign(n * 32, 0);
  if (jobs.empty()) { sc->unit_digests.assign(std::max<size_t>(n * 32, 1), 0); return; }
  cudaStream_t st = ctx->stream;
  ctx->query_in.reserve(sizeof(HashJob) * jobs.size());
  ctx->query_out.reserve(32 * jobs.size());
  CK(cudaMemcpyAsync(ctx->query_in.p, jobs.data(), sizeof(HashJob) * jobs.size(), cudaMemcpyHostToDevice, st));
  k_sha256<<<(unsigned)((jobs.size() + 31) / 32), 32, 0, st>>>(sc->chunk_blob, ctx->query_in.as<HashJob>(), (u32)jobs.size(), ctx->query_out.as<u8>());
  CK(cudaGetLastError());
  ctx->launches++;
  std::vector<uint8_t> dig(32 * jobs.size());
  CK(cudaMemcpyAsync(dig.data(), ctx->query_out.p, dig.size(), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (size_t i = 0; i < jobs.size(); i++) memcpy(&sc->unit_digests[(size_t)job_unit[i] * 32], &dig[i * 32], 32);
}

struct Assembler {
  dcdf_ctx* ctx;
  const dcdf_superchunk* sc;
  uint32_t slice;
  dcdf_saved* out;
  std::map<CidBytes, size_t> seen;   // objects already in out->nodes
  std::vector<uint8_t> dacs;         // host copy of the slice's Dac bytes
  uint64_t dac_base = 0;

  void add(SavedNode&& n) {
    if (seen.count(n.cid)) return;
    seen[n.cid] = out->nodes.size();
    out->nodes.push_back(std::move(n));
  }

  struct Result {
    std::vector<uint8_t> body;  // NODE_SUPERCHUNK byte + Superchunk::save_to
    uint64_t data_size;         // Superchunk::size()  superchunk.rs:654-669 (as written there: the encoding byte is not counted)
    dcdf_build_stats stats;     // MMStruct3Build of this node (superchunk.rs:261-269)
  };

  Result assemble(uint32_t node) {
    const uint32_t n_nodes = (uint32_t)sc->nodes.size();
    const TreeNode& nd = sc->nodes[node];
    const auto& g = sc->geom[node];
    const auto& sl = sc->slices[slice];
    const NodeState& ns = sc->nstate[(size_t)slice * n_nodes + node];
    const bool is_float = sc->encoding == DCDF_ENC_F32 || sc->encoding == DCDF_ENC_F64;
    Result R;
    memset(&R.stats, 0, sizeof R.stats);
    std::vector<uint8_t> refs;
    std::vector<CidBytes> links;
    std::map<CidBytes, uint32_t> link_index;  // external_references: HashMap<Cid, usize>  superchunk.rs:199,222-232
    uint64_t sizes = 0;
    for (u32 c = 0; c < nd.n_children; c++) {
      const TreeChild& ch = sc->children[nd.first_child + c];
      CidBytes cid;
      bool stored = false;
      if (ch.kind == 1) {
        const uint32_t u = sl.unit_base + (uint32_t)ch.index;
        if (sc->stored[u]) {
          stored = true;
          cid = cid_of_digest(&sc->unit_digests[(size_t)u * 32]);
          SavedNode sn;
          sn.cid = cid; sn.type = NODE_SUBCHUNK;
          put_header(sn.bytes, NODE_MMSTRUCT3);
          sn.bytes.push_back(NODE_SUBCHUNK);
          sn.dev_off = sc->chunk_off[u]; sn.dev_len = sc->results[u].bytes;
          add(std::move(sn));
          sizes += sc->results[u].bytes + 1;          // build.data.size(): MMStruct3::size = Chunk::size + 1 (mmstruct.rs:186-196)
          R.stats.snapshots += sc->results[u].snapshots;
          R.stats.logs += sc->results[u].logs;
        }
      } else if (ch.kind == 2 && sc->nstate[(size_t)slice * n_nodes + ch.index].alive) {
        stored = true;
        Result sub = assemble((uint32_t)ch.index);    // the recursion saves its own subchunks and Links first (superchunk.rs:171)
        SavedNode sn;
        sn.type = NODE_SUPERCHUNK;
        put_header(sn.bytes, NODE_MMSTRUCT3);
        sn.bytes.insert(sn.bytes.end(), sub.body.begin(), sub.body.end());
        sn.cid = cid_of_bytes(sn.bytes);
        cid = sn.cid;
        add(std::move(sn));
        sizes += sub.data_size + 1;
        R.stats.snapshots += sub.stats.snapshots;
        R.stats.logs += sub.stats.logs;
      }
      if (!stored) {
        R.stats.elided++;
        refs.push_back(0);                            // REFERENCE_ELIDED  superchunk.rs:833
        continue;
      }
      auto it = link_index.find(cid);
      uint32_t index;
      if (it == link_index.end()) {
        index = (uint32_t)links.size();
        links.push_back(cid);
        link_index[cid] = index;
      } else {
        index = it->second;
      }
      refs.push_back(2);                              // REFERENCE_EXTERNAL
      put_u32(refs, index);
    }
    // Links node (links.rs:65-76) and its CID
    SavedNode ln;
    ln.type = NODE_LINKS;
    put_header(ln.bytes, NODE_LINKS);
    put_u32(ln.bytes, (uint32_t)links.size());
    for (auto& c : links) ln.bytes.insert(ln.bytes.end(), c.begin(), c.end());
    ln.cid = cid_of_bytes(ln.bytes);
    const CidBytes external_cid = ln.cid;
    const uint64_t size_external = 7 + 4 + (uint64_t)DCDF_CID_BYTES * links.size();  // Links::size  links.rs:96-100
    add(std::move(ln));
    // Superchunk::save_to  superchunk.rs:683-707
    const size_t di = ((size_t)slice * n_nodes + node) * 2;
    const uint64_t max_off = sc->node_dac_off[di] - dac_base, max_len = sc->node_dac_size[di];
    const uint64_t min_off = sc->node_dac_off[di + 1] - dac_base, min_len = sc->node_dac_size[di + 1];
    std::vector<uint8_t>& b = R.body;
    b.push_back(NODE_SUPERCHUNK);
    put_u32(b, (uint32_t)sl.info.shape[0]); put_u32(b, (uint32_t)g.rows); put_u32(b, (uint32_t)g.cols);
    put_u32(b, (uint32_t)g.sidelen);
    b.push_back((uint8_t)g.levels);
    put_u32(b, (uint32_t)g.chunks_sidelen);
    put_u32(b, (uint32_t)g.subsidelen);
    b.push_back((uint8_t)(is_float ? ns.bits : 0));
    b.push_back((uint8_t)sc->encoding);
    put_u32(b, nd.n_children);
    const size_t refs_bytes = refs.size();
    b.insert(b.end(), refs.begin(), refs.end());
    b.insert(b.end(), external_cid.begin(), external_cid.end());
    put_u32(b, 0);  // n_local: build never produces Local references
    b.insert(b.end(), dacs.begin() + max_off, dacs.begin() + max_off + max_len);
    b.insert(b.end(), dacs.begin() + min_off, dacs.begin() + min_off + min_len);
    R.data_size = 7 + 4 * 3 + 4 + 1 + 4 + 4 + 1 + 4 + refs_bytes + DCDF_CID_BYTES + 4 + max_len + min_len;
    R.stats.external = (uint32_t)links.size();
    R.stats.local = 0;
    R.stats.size = R.data_size + size_external + sizes;  // superchunk.rs:263
    return R;
  }
};

}  // namespace

extern "C" {

int32_t dcdf_superchunk_save(dcdf_ctx* ctx, const dcdf_superchunk* sc, uint32_t slice, dcdf_saved** out) {
  return guarded(ctx, [&] {
    if (!out) api_fail(DCDF_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!sc || slice >= sc->slices.size()) api_fail(DCDF_ERR_BAD_ARG, "bad slice");
    ensure_digests(ctx, sc);
    const uint32_t n_nodes = (uint32_t)sc->nodes.size();
    std::unique_ptr<dcdf_saved> sv(new dcdf_saved());
    sv->device = ctx->device;
    sv->sc = sc;
    Assembler A{ctx, sc, slice, sv.get(), {}, {}, 0};
    // the slice's Dac bytes (all nodes, max then min) are contiguous in the device blob
    const size_t d0 = (size_t)slice * n_nodes * 2, d1 = d0 + (size_t)n_nodes * 2;
    A.dac_base = sc->node_dac_off[d0];
    const uint64_t dac_end = sc->node_dac_off[d1 - 1] + sc->node_dac_size[d1 - 1];
    A.dacs.resize((size_t)(dac_end - A.dac_base) + 1);
    if (dac_end > A.dac_base) {
      CK(cudaMemcpyAsync(A.dacs.data(), sc->dac_blob + A.dac_base, (size_t)(dac_end - A.dac_base), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    Assembler::Result root = A.assemble(0);
    SavedNode rn;
    rn.type = NODE_SUPERCHUNK;
    put_header(rn.bytes, NODE_MMSTRUCT3);
    rn.bytes.insert(rn.bytes.end(), root.body.begin(), root.body.end());
    rn.cid = cid_of_bytes(rn.bytes);
    sv->nodes.push_back(std::move(rn));  // always last, even if an identical node is already listed
    sv->stats = root.stats;
    *out = sv.release();
  });
}

int32_t dcdf_saved_free(dcdf_saved* s) {
  delete s;
  return DCDF_OK;
}

int32_t dcdf_saved_count(const dcdf_saved* s, uint32_t* n_nodes) {
  if (!s || !n_nodes) return DCDF_ERR_BAD_ARG;
  *n_nodes = (uint32_t)s->nodes.size();
  return DCDF_OK;
}

int32_t dcdf_saved_node(const dcdf_saved* s, uint32_t i, uint8_t* cid, int32_t* node_type, uint64_t* size) {
  if (!s || i >= s->nodes.size()) return DCDF_ERR_BAD_ARG;
  const SavedNode& n = s->nodes[i];
  if (cid) memcpy(cid, n.cid.data(), DCDF_CID_BYTES);
  if (node_type) *node_type = n.type;
  if (size) *size = n.bytes.size() + n.dev_len;
  return DCDF_OK;
}

int32_t dcdf_saved_node_bytes(dcdf_ctx* ctx, const dcdf_saved* s, uint32_t i, uint8_t* dst, uint64_t cap, int32_t mem) {
  return guarded(ctx, [&] {
    if (!s || i >= s->nodes.size() || !dst) api_fail(DCDF_ERR_BAD_ARG, "bad node index / null destination");
    const SavedNode& n = s->nodes[i];
    const uint64_t total = n.bytes.size() + n.dev_len;
    if (cap < total) api_fail(DCDF_ERR_BAD_ARG, "destination too small: need %llu bytes", (unsigned long long)total);
    const bool dev = mem == DCDF_MEM_DEVICE;
    CK(cudaMemcpyAsync(dst, n.bytes.data(), n.bytes.size(), dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost, ctx->stream));
    if (n.dev_len)
      CK(cudaMemcpyAsync(dst + n.bytes.size(), s->sc->chunk_blob + n.dev_off, n.dev_len, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                         ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  });
}

int32_t dcdf_saved_all_bytes(dcdf_ctx* ctx, const dcdf_saved* s, uint8_t* dst, uint64_t cap, uint64_t* offsets) {
  return guarded(ctx, [&] {
    if (!s || !offsets) api_fail(DCDF_ERR_BAD_ARG, "null saved object / offsets");
    uint64_t total = 0;
    for (size_t i = 0; i < s->nodes.size(); i++) {
      const SavedNode& n = s->nodes[i];
      offsets[i] = total;
      total += n.bytes.size() + n.dev_len;
    }
    offsets[s->nodes.size()] = total;
    if (!dst) return;
    if (cap < total) api_fail(DCDF_ERR_BAD_ARG, "destination too small: need %llu bytes", (unsigned long long)total);
    // every device-resident part is queued without waiting in between; one wait at the end
    for (size_t i = 0; i < s->nodes.size(); i++) {
      const SavedNode& n = s->nodes[i];
      uint8_t* d = dst + offsets[i];
      if (!n.bytes.empty()) memcpy(d, n.bytes.data(), n.bytes.size());
      if (n.dev_len) CK(cudaMemcpyAsync(d + n.bytes.size(), s->sc->chunk_blob + n.dev_off, n.dev_len, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
  });
}

int32_t dcdf_saved_stats(const dcdf_saved* s, dcdf_build_stats* stats) {
  if (!s || !stats) return DCDF_ERR_BAD_ARG;
  *stats = s->stats;
  return DCDF_OK;
}

}  // extern "C"

// ===================================================================================== open from stored bytes
namespace {

struct Rd {
  const uint8_t* p;
  uint64_t n, pos = 0;
  void need(uint64_t k) const { if (pos + k > n) api_fail(DCDF_ERR_BAD_FORMAT, "stored node ends early"); }
  uint8_t u8_() { need(1); return p[pos++]; }
  uint32_t u32_() { need(4); uint32_t w = ((uint32_t)p[pos] << 24) | ((uint32_t)p[pos + 1] << 16) | ((uint32_t)p[pos + 2] << 8) | p[pos + 3]; pos += 4; return w; }
  const uint8_t* take(uint64_t k) { need(k); const uint8_t* q = p + pos; pos += k; return q; }
};

uint8_t read_header(Rd& r) {  // resolver.rs:160-172
  if (r.u8_() != 0xDC || r.u8_() != 0xE0) api_fail(DCDF_ERR_BAD_FORMAT, "not a DCDF graph node (magic number)");
  if (r.u32_() != 1) api_fail(DCDF_ERR_BAD_FORMAT, "unrecognized node format version");
  return r.u8_();
}

// A Dac as a byte range plus decoded values (dac.rs:48-63, :80-93)
struct HostDac {
  const uint8_t* bytes = nullptr;
  uint64_t size = 0;
  struct Level { uint32_t len; const uint8_t* index; const uint8_t* words; const uint8_t* data; };
  std::vector<Level> levels;
  static uint32_t be(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
  void parse(Rd& r) {
    const uint64_t start = r.pos;
    bytes = r.p + start;
    const uint32_t nl = r.u8_();
    if (nl > 8) api_fail(DCDF_ERR_BAD_FORMAT, "Dac with more than 8 levels");
    for (uint32_t j = 0; j < nl; j++) {
      Level L;
      L.len = r.u32_();
      if (r.u32_() != 4) api_fail(DCDF_ERR_BAD_FORMAT, "BitMap k != 4");
      L.index = r.take(4ull * (L.len / 128));
      L.words = r.take(4ull * ((L.len + 31) / 32));
      L.data = r.take(L.len);
      if (j > 0 && L.len > levels[j - 1].len) api_fail(DCDF_ERR_BAD_FORMAT, "Dac level longer than the one below");
      levels.push_back(L);
    }
    size = r.pos - start;
  }
  uint64_t len() const { return levels.empty() ? 0 : levels[0].len; }
  static bool bit(const Level& L, uint32_t i) { return (L.words[i >> 3] >> (7 - (i & 7))) & 1; }
  static uint32_t rank(const Level& L, uint32_t i) {  // bitmap.rs:186-212
    uint32_t block = i / 128, count = block ? be(L.index + 4 * (block - 1)) : 0;
    for (uint32_t w = block * 4; w < i / 32; w++) count += (uint32_t)__builtin_popcount(be(L.words + 4 * w));
    if (i % 32) count += (uint32_t)__builtin_popcount(be(L.words + 4 * (i / 32)) >> (32 - i % 32));
    return count;
  }
  int64_t get(uint32_t index) const {
    uint64_t v = 0;
    for (size_t j = 0; j < levels.size(); j++) {
      const Level& L = levels[j];
      if (index >= L.len) api_fail(DCDF_ERR_BAD_FORMAT, "Dac continuation bits point past the next level");
      v |= (uint64_t)L.data[index] << (8 * j);
      if (!bit(L, index)) break;
      index = rank(L, index);
    }
    return (int64_t)((v >> 1) ^ (0 - (v & 1)));
  }
};

struct PNode {
  int64_t shape[3];
  int64_t sidelen, chunks_sidelen, subsidelen;
  uint32_t levels;
  int bits, encoding;
  struct Ref { int kind; const uint8_t* chunk = nullptr; uint64_t chunk_len = 0; int child = -1; };
  std::vector<Ref> refs;
  HostDac max, min;
};

struct Loader {
  dcdf_fetch_fn fetch;
  void* user;
  std::vector<PNode> pool;

  Rd get(const uint8_t* cid) {
    const uint8_t* bytes = nullptr;
    uint64_t len = 0;
    if (fetch(user, cid, &bytes, &len) != 0 || !bytes) api_fail(DCDF_ERR_BAD_ARG, "the fetch callback could not provide a node");
    return Rd{bytes, len};
  }
  // body of a superchunk node (after the NODE_SUPERCHUNK byte)  superchunk.rs:713-768
  int parse_super(Rd& r, int depth) {
    if (depth > 4) api_fail(DCDF_ERR_BAD_ARG, "superchunks nested more than four levels deep are not supported");
    PNode n;
    for (int i = 0; i < 3; i++) n.shape[i] = r.u32_();
    n.sidelen = r.u32_();
    n.levels = r.u8_();
    n.chunks_sidelen = r.u32_();
    n.subsidelen = r.u32_();
    n.bits = r.u8_();
    n.encoding = r.u8_();
    if (!(n.encoding == 4 || n.encoding == 8 || n.encoding == 32 || n.encoding == 64)) api_fail(DCDF_ERR_BAD_FORMAT, "bad encoding byte");
    const uint32_t n_refs = r.u32_();
    if (n.subsidelen <= 0 || n.subsidelen > 4096 || (uint64_t)n_refs != (uint64_t)n.subsidelen * (uint64_t)n.subsidelen ||
        n.chunks_sidelen <= 0 || n.chunks_sidelen * n.subsidelen != n.sidelen)
      api_fail(DCDF_ERR_BAD_FORMAT, "superchunk geometry does not add up");
    std::vector<std::pair<int, uint32_t>> raw(n_refs);
    for (auto& x : raw) {
      x.first = r.u8_();
      x.second = 0;
      if (x.first == 1 || x.first == 2) x.second = r.u32_();
      else if (x.first != 0) api_fail(DCDF_ERR_BAD_FORMAT, "unrecognized reference type");
      if (x.first == 1) api_fail(DCDF_ERR_BAD_ARG, "Local references are never produced by Superchunk::build and are not supported");
    }
    const uint8_t* ext = r.take(DCDF_CID_BYTES);
    if (ext[0] != 1 || ext[2] != 0x12 || ext[3] != 0x20) api_fail(DCDF_ERR_BAD_ARG, "only CIDv1 with a SHA2-256 multihash is supported");
    if (r.u32_() != 0) api_fail(DCDF_ERR_BAD_ARG, "Local subchunks are not supported");
    n.max.parse(r);
    n.min.parse(r);
    if (r.pos != r.n) api_fail(DCDF_ERR_BAD_FORMAT, "trailing bytes after the superchunk node");
    const uint64_t want = (uint64_t)n_refs * (uint64_t)n.shape[0];
    if (n.max.len() != want || n.min.len() != want) api_fail(DCDF_ERR_BAD_FORMAT, "min / max Dac length does not match instants x subchunks");
    // Links (links.rs:41-63)
    std::vector<const uint8_t*> links;
    bool any_ext = false;
    for (auto& x : raw) any_ext = any_ext || x.first == 2;
    if (any_ext) {
      Rd lr = get(ext);
      if (read_header(lr) != NODE_LINKS) api_fail(DCDF_ERR_BAD_FORMAT, "expecting a Links node");
      const uint32_t nl = lr.u32_();
      for (uint32_t i = 0; i < nl; i++) links.push_back(lr.take(DCDF_CID_BYTES));
    }
    n.refs.resize(n_refs);
    const int self = (int)pool.size();
    pool.push_back(PNode());
    for (uint32_t i = 0; i < n_refs; i++) {
      PNode::Ref& ref = n.refs[i];
      ref.kind = raw[i].first;
      if (ref.kind != 2) continue;
      if (raw[i].second >= links.size()) api_fail(DCDF_ERR_BAD_FORMAT, "External reference past the end of Links");
      Rd cr = get(links[raw[i].second]);
      if (read_header(cr) != NODE_MMSTRUCT3) api_fail(DCDF_ERR_BAD_FORMAT, "expecting an MMStruct3 node");
      const uint8_t t = cr.u8_();
      if (t == NODE_SUBCHUNK) {
        ref.chunk = cr.p + cr.pos;
        ref.chunk_len = cr.n - cr.pos;
        if (ref.chunk_len < 6 || ref.chunk_len > 0xfffffff0ull) api_fail(DCDF_ERR_BAD_FORMAT, "bad Chunk size");
      } else if (t == NODE_SUPERCHUNK) {
        Rd body{cr.p + cr.pos, cr.n - cr.pos};
        ref.child = parse_super(body, depth + 1);
      } else {
        api_fail(DCDF_ERR_BAD_ARG, "only Subchunk and Superchunk nodes can be opened (Span nodes stay on the host side)");
      }
    }
    pool[self] = std::move(n);
    return self;
  }
};

}  // namespace

extern "C" int32_t dcdf_superchunk_open(dcdf_ctx* ctx, uint32_t n_slices, const uint8_t* root_cids, dcdf_fetch_fn fetch, void* user,
                                        dcdf_superchunk** out) {
  return guarded(ctx, [&] {
    if (!out) api_fail(DCDF_ERR_BAD_ARG, "null out");
    *out = nullptr;
    if (!n_slices || !root_cids || !fetch) api_fail(DCDF_ERR_BAD_ARG, "null argument");
    Loader L{fetch, user, {}};
    std::vector<int> roots;
    for (uint32_t s = 0; s < n_slices; s++) {
      Rd r = L.get(root_cids + (size_t)s * DCDF_CID_BYTES);
      if (read_header(r) != NODE_MMSTRUCT3) api_fail(DCDF_ERR_BAD_FORMAT, "expecting an MMStruct3 node");
      if (r.u8_() != NODE_SUPERCHUNK) api_fail(DCDF_ERR_BAD_ARG, "the root of slice %u is not a Superchunk node", s);
      Rd body{r.p + r.pos, r.n - r.pos};
      roots.push_back(L.parse_super(body, 0));
    }
    // ---- k2_levels from the stored nodes: one entry per depth (every node of a depth must agree) + the leaf level
    const PNode& r0 = L.pool[roots[0]];
    const int64_t rows = r0.shape[1], cols = r0.shape[2];
    std::vector<uint32_t> levels;
    std::vector<int64_t> leaf_cs;
    std::function<void(int, size_t)> walk = [&](int pn, size_t depth) {
      const PNode& n = L.pool[pn];
      if (levels.size() <= depth) { levels.resize(depth + 1, 0xffffffffu); leaf_cs.resize(depth + 1, 0); }
      if (levels[depth] == 0xffffffffu) { levels[depth] = n.levels; leaf_cs[depth] = n.chunks_sidelen; }
      else if (levels[depth] != n.levels || leaf_cs[depth] != n.chunks_sidelen) api_fail(DCDF_ERR_BAD_ARG, "superchunk nodes of one depth disagree on their geometry");
      for (auto& ref : n.refs) if (ref.child >= 0) walk(ref.child, depth + 1);
    };
    int64_t T_total = 0;
    for (uint32_t s = 0; s < n_slices; s++) {
      const PNode& rn = L.pool[roots[s]];
      if (rn.shape[1] != rows || rn.shape[2] != cols || rn.encoding != r0.encoding || rn.sidelen != r0.sidelen)
        api_fail(DCDF_ERR_BAD_ARG, "the slices of one array must share shape, sidelen and encoding");
      if (s + 1 < n_slices && rn.shape[0] != r0.shape[0]) api_fail(DCDF_ERR_BAD_ARG, "all slices but the last must have the same number of instants");
      if (rn.shape[0] <= 0 || rn.shape[0] > r0.shape[0]) api_fail(DCDF_ERR_BAD_ARG, "bad number of instants in slice %u", s);
      walk(roots[s], 0);
      T_total += rn.shape[0];
    }
    const int64_t leaf_side = leaf_cs.back();
    uint32_t leaf_levels = 0;
    while (((int64_t)1 << leaf_levels) < leaf_side) leaf_levels++;
    if (((int64_t)1 << leaf_levels) != leaf_side) api_fail(DCDF_ERR_BAD_FORMAT, "chunks_sidelen is not a power of two");
    levels.push_back(leaf_levels);
    uint32_t sum = 0;
    for (uint32_t l : levels) sum += l;
    if (sum != levels_for(std::max(rows, cols), 2) || ((int64_t)1 << sum) != r0.sidelen) api_fail(DCDF_ERR_BAD_FORMAT, "stored levels do not match the stored shape");
    TreeGeo G;
    build_tree(G, rows, cols, levels.data(), (uint32_t)levels.size());
    const uint32_t n_nodes = (uint32_t)G.nodes.size();
    const uint32_t n_slots = (uint32_t)(G.leaf_grid * G.leaf_grid);
    const int ls = G.leaf_side;

    std::unique_ptr<dcdf_superchunk> sc(new dcdf_superchunk());
    sc->device = ctx->device;
    sc->opened = true;
    sc->encoding = r0.encoding;
    sc->shape[0] = T_total; sc->shape[1] = rows; sc->shape[2] = cols;
    sc->chunk_size = r0.shape[0];
    sc->n_slots = n_slots;
    sc->nodes = G.nodes; sc->children = G.children; sc->geom = G.geom; sc->leaf_unit = G.leaf_unit;
    sc->leaf_rows = G.leaf_rows; sc->leaf_cols = G.leaf_cols; sc->leaf_side = G.leaf_side; sc->leaf_grid = G.leaf_grid;
    sc->tbl_per_instant = G.tbl_per_instant;
    sc->slot_unit.assign((size_t)n_slices * n_slots, -1);
    sc->nstate.assign((size_t)n_slices * n_nodes, NodeState{0, 0});
    sc->node_dac_off.assign((size_t)n_slices * n_nodes * 2, 0);
    sc->node_dac_size.assign((size_t)n_slices * n_nodes * 2, 0);
    std::vector<uint8_t> blob, dac_blob;
    std::vector<int64_t> tbl_max, tbl_min;
    uint64_t dir_total = 0;
    for (uint32_t s = 0; s < n_slices; s++) {
      const PNode& rn = L.pool[roots[s]];
      const int inst = (int)rn.shape[0];
      dcdf_superchunk::Slice sl;
      memset(&sl, 0, sizeof sl);
      sl.t0 = (int64_t)s * sc->chunk_size;
      sl.unit_base = (uint32_t)sc->units.size();
      sl.table_base = tbl_max.size();
      sl.info.shape[0] = inst;
      tbl_max.resize(tbl_max.size() + (size_t)inst * G.tbl_per_instant, 0);
      tbl_min.resize(tbl_min.size() + (size_t)inst * G.tbl_per_instant, 0);
      for (int gr = 0; gr < G.leaf_rows; gr++)
        for (int gc = 0; gc < G.leaf_cols; gc++) {
          EncUnit u;
          memset(&u, 0, sizeof u);
          u.rows = (int)std::min<int64_t>(ls, rows - (int64_t)gr * ls); u.cols = (int)std::min<int64_t>(ls, cols - (int64_t)gc * ls);
          u.instants = inst;
          u.slot = (uint32_t)((int64_t)gr * G.leaf_grid + gc);
          u.row0 = gr * ls; u.col0 = gc * ls;
          u.piece_base = (uint32_t)dir_total;
          dir_total += (uint64_t)inst;
          sc->slot_unit[(size_t)s * n_slots + u.slot] = (int32_t)sc->units.size();
          sc->units.push_back(u);
          sc->stored.push_back(0);
          sc->chunk_off.push_back(0);
          sc->results.push_back(UnitResult{0, 0, 0});
        }
      if (dir_total > 0xfffffff0ull) api_fail(DCDF_ERR_BAD_ARG, "too many (chunk, instant) pairs for one handle");
      sl.n_units = (uint32_t)sc->units.size() - sl.unit_base;
      sl.chunk_blob_off = blob.size();
      // stored node tree against the static tree
      std::function<void(int, uint32_t)> place = [&](int pn, uint32_t node) {
        const PNode& n = L.pool[pn];
        const TreeNode& nd = G.nodes[node];
        const auto& g = G.geom[node];
        if (n.shape[0] != inst || n.shape[1] != g.rows || n.shape[2] != g.cols || n.sidelen != g.sidelen || n.levels != g.levels ||
            n.chunks_sidelen != g.chunks_sidelen || n.subsidelen != g.subsidelen || n.refs.size() != nd.n_children || n.encoding != r0.encoding)
          api_fail(DCDF_ERR_BAD_FORMAT, "stored superchunk node does not match the geometry of its position");
        sc->nstate[(size_t)s * n_nodes + node] = NodeState{1, n.bits};
        const size_t di = ((size_t)s * n_nodes + node) * 2;
        sc->node_dac_off[di] = dac_blob.size(); sc->node_dac_size[di] = n.max.size;
        dac_blob.insert(dac_blob.end(), n.max.bytes, n.max.bytes + n.max.size);
        sc->node_dac_off[di + 1] = dac_blob.size(); sc->node_dac_size[di + 1] = n.min.size;
        dac_blob.insert(dac_blob.end(), n.min.bytes, n.min.bytes + n.min.size);
        int64_t* tmax = tbl_max.data() + sl.table_base + (size_t)nd.tbl_off * inst;
        int64_t* tmin = tbl_min.data() + sl.table_base + (size_t)nd.tbl_off * inst;
        const uint64_t cells = (uint64_t)inst * nd.n_children;
        for (uint64_t i = 0; i < cells; i++) { tmax[i] = n.max.get((uint32_t)i); tmin[i] = n.min.get((uint32_t)i); }
        for (u32 c = 0; c < nd.n_children; c++) {
          const TreeChild& ch = G.children[nd.first_child + c];
          const PNode::Ref& ref = n.refs[c];
          if (ref.kind == 0) continue;
          if (ch.kind == 0) api_fail(DCDF_ERR_BAD_FORMAT, "a reference outside the raster is not Elided");
          if (ch.kind == 1) {
            if (!ref.chunk) api_fail(DCDF_ERR_BAD_FORMAT, "expecting a Subchunk node at a leaf position");
            const uint32_t u = sl.unit_base + (uint32_t)ch.index;
            sc->stored[u] = 1;
            sc->chunk_off[u] = blob.size();
            sc->results[u].bytes = ref.chunk_len;
            sc->units[u].bits = ref.chunk[1];
            blob.insert(blob.end(), ref.chunk, ref.chunk + ref.chunk_len);
            blob.resize((blob.size() + 15) & ~size_t(15), 0);  // chunks start 16-byte aligned, like the gather kernel lays them out
          } else {
            if (ref.child < 0) api_fail(DCDF_ERR_BAD_FORMAT, "expecting a Superchunk node at an inner position");
            place(ref.child, (uint32_t)ch.index);
          }
        }
      };
      place(roots[s], 0);
      sl.info.chunk_bytes = blob.size() - sl.chunk_blob_off;
      sc->slices.push_back(sl);
    }
    sc->chunk_off.push_back(blob.size());
    // ---- upload
    cudaStream_t st = ctx->stream;
    sc->chunk_blob_size = blob.size();
    sc->chunk_blob = static_cast<uint8_t*>(pool_alloc(blob.size() + 64, st));
    CK(cudaMemsetAsync(sc->chunk_blob + blob.size(), 0, 64, st));
    if (!blob.empty()) CK(cudaMemcpyAsync(sc->chunk_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
    sc->dac_blob_size = dac_blob.size();
    sc->dac_blob = static_cast<uint8_t*>(pool_alloc(dac_blob.size() + 16, st));
    if (!dac_blob.empty()) CK(cudaMemcpyAsync(sc->dac_blob, dac_blob.data(), dac_blob.size(), cudaMemcpyHostToDevice, st));
    sc->tbl_len = tbl_max.size();
    sc->tbl_max = static_cast<int64_t*>(pool_alloc(sizeof(int64_t) * std::max<size_t>(tbl_max.size(), 1), st));
    sc->tbl_min = static_cast<int64_t*>(pool_alloc(sizeof(int64_t) * std::max<size_t>(tbl_min.size(), 1), st));
    if (!tbl_max.empty()) {
      CK(cudaMemcpyAsync(sc->tbl_max, tbl_max.data(), sizeof(int64_t) * tbl_max.size(), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(sc->tbl_min, tbl_min.data(), sizeof(int64_t) * tbl_min.size(), cudaMemcpyHostToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
    dcdf_superchunk* raw = sc.release();
    try {
      build_super_meta(ctx, raw);  // Chunk::read_from's validation for every stored chunk, now rather than at the first query
    } catch (...) {
      dcdf_superchunk_free(raw);
      throw;
    }
    *out = raw;
  });
}
