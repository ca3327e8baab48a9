/*
 * oracle/ipc_writer.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Uncompressed Arrow IPC *stream* writer for the oracle's ArrayData model: what
 * ipc.NewWriter(buf, WithSchema, WithAllocator) + Write + Close produce in the reference's
 * offline V2 path (reporter/parca_reporter.go:1779-1790). arrow-go's ipc package and
 * google/flatbuffers v25.12.19 (go.mod:12, :102) are not vendored; this follows the published
 * Arrow IPC format (format/Message.fbs, format/Schema.fbs, "Encapsulated message format"):
 *   - messages: 0xFFFFFFFF, int32 metadata length (padded so the body starts 8-byte aligned),
 *     flatbuffer Message, body; stream ends with 0xFFFFFFFF 0x00000000
 *   - dictionary ids assigned in schema pre-order; dictionary batches emitted inner-first,
 *     all before the record batch
 *   - body buffers 8-byte aligned, Buffer.length = unpadded size, a zero-length validity
 *     buffer when null_count == 0, RunEndEncoded parents carry no buffers
 * Flatbuffer vtable/field placement inside the metadata is this writer's own (any conforming
 * reader accepts it); byte-level agreement with a Go build is UNPINNED.
 */
#ifndef ORACLE_IPC_WRITER_H
#define ORACLE_IPC_WRITER_H
#include <cassert>
#include "arrow_model.h"

namespace orc {

// ---- minimal flatbuffer builder (back-to-front, as the format requires) ---------------------
class FBB {
 public:
  typedef uint32_t Off;
  FBB() : buf_(1024), head_(1024) {}
  uint32_t size() const { return (uint32_t)(buf_.size() - head_); }
  const uint8_t* data() const { return buf_.data() + head_; }

  Off string(std::string_view s) {
    prep(4, s.size() + 1);
    push_bytes("\0", 1);
    push_bytes(s.data(), s.size());
    push<uint32_t>((uint32_t)s.size());
    return size();
  }
  Off vec_offsets(const std::vector<Off>& v) {
    prep(4, v.size() * 4);
    for (size_t i = v.size(); i-- > 0;) push<uint32_t>(refer(v[i]));
    push<uint32_t>((uint32_t)v.size());
    return size();
  }
  // vector of 16-byte structs {int64,int64} (FieldNode / Buffer), or of int64 (w=1)
  Off vec_i64(const std::vector<int64_t>& flat, int per_elem) {
    prep(4, flat.size() * 8);
    prep(8, flat.size() * 8);
    for (size_t i = flat.size(); i-- > 0;) push<int64_t>(flat[i]);
    push<uint32_t>((uint32_t)(flat.size() / per_elem));
    return size();
  }
  void start(int nslots) { slots_.assign(nslots, 0); obj_start_ = size(); }
  template <class T>
  void scalar(int slot, T v, T dflt) {
    if (v == dflt) return;
    prep(sizeof(T), 0);
    push<T>(v);
    slots_[slot] = size();
  }
  void offset(int slot, Off o) {
    if (!o) return;
    prep(4, 0);
    push<uint32_t>(refer(o));
    slots_[slot] = size();
  }
  Off end() {
    prep(4, 0);
    push<int32_t>(0);
    Off tab = size();
    int n = (int)slots_.size();
    while (n > 0 && slots_[n - 1] == 0) n--;
    for (int i = n; i-- > 0;) push<uint16_t>(slots_[i] ? (uint16_t)(tab - slots_[i]) : 0);
    push<uint16_t>((uint16_t)(tab - obj_start_));
    push<uint16_t>((uint16_t)((n + 2) * 2));
    int32_t so = (int32_t)size() - (int32_t)tab;
    memcpy(buf_.data() + buf_.size() - tab, &so, 4);
    return tab;
  }
  void finish(Off root) {
    prep(minalign_, 4);
    push<uint32_t>(refer(root));
  }

 private:
  std::vector<uint8_t> buf_;
  size_t head_;
  size_t minalign_ = 1;
  std::vector<uint32_t> slots_;
  uint32_t obj_start_ = 0;
  void grow(size_t need) {
    while (head_ < need) {
      size_t old = buf_.size();
      std::vector<uint8_t> nb(old * 2);
      memcpy(nb.data() + old + head_, buf_.data() + head_, old - head_);
      head_ += old;
      buf_.swap(nb);
    }
  }
  void prep(size_t align, size_t extra) {
    if (align > minalign_) minalign_ = align;
    size_t pad = (~(size() + extra) + 1) & (align - 1);
    grow(pad + extra + align + 8);
    for (size_t i = 0; i < pad; i++) buf_[--head_] = 0;
  }
  template <class T>
  void push(T v) { grow(sizeof(T)); head_ -= sizeof(T); memcpy(buf_.data() + head_, &v, sizeof(T)); }
  void push_bytes(const char* p, size_t n) { grow(n); head_ -= n; if (n) memcpy(buf_.data() + head_, p, n); }
  uint32_t refer(Off o) { prep(4, 0); return size() - o + 4; }
};

// ---- Arrow flatbuffer enums -------------------------------------------------------------
enum FbType { FB_Int = 2, FB_Binary = 4, FB_Bool = 6, FB_List = 12, FB_Utf8 = 5, FB_Timestamp = 10, FB_Struct = 13, FB_FSB = 15, FB_REE = 22, FB_Utf8View = 24, FB_ListView = 25 };
enum FbHeader { FB_Schema = 1, FB_DictionaryBatch = 2, FB_RecordBatch = 3 };

class IpcWriter {
 public:
  std::vector<uint8_t> out;

  void write_stream(const std::vector<Field>& fields, const std::vector<std::pair<std::string, std::string>>& schema_md,
                    const std::vector<ArrayData>& cols, int64_t nrows) {
    out.clear();
    next_id_ = 0;
    // schema message (dictionary ids in pre-order)
    {
      FBB b;
      std::vector<FBB::Off> fo;
      for (auto& f : fields) fo.push_back(field(b, f));
      FBB::Off fv = b.vec_offsets(fo);
      FBB::Off md = kv(b, schema_md);
      b.start(4);
      b.offset(1, fv);
      b.offset(2, md);
      FBB::Off sch = b.end();
      message(b, FB_Schema, sch, 0);
      emit(b, nullptr, 0);
    }
    // dictionaries, inner-first, ids re-derived by the same pre-order walk
    int64_t id = 0;
    for (size_t i = 0; i < fields.size(); i++) dicts(fields[i].type, cols[i], id);
    // record batch
    {
      Body body;
      for (auto& c : cols) encode(c, body);
      FBB b;
      FBB::Off rb = record_batch(b, nrows, body);
      message(b, FB_RecordBatch, rb, (int64_t)body.bytes.size());
      emit(b, body.bytes.data(), body.bytes.size());
    }
    // end of stream
    uint32_t eos[2] = {0xFFFFFFFFu, 0};
    append(eos, 8);
  }

 private:
  struct Body {
    std::vector<int64_t> nodes;    // length, null_count pairs
    std::vector<int64_t> buffers;  // offset, length pairs
    std::vector<int64_t> variadic;
    std::vector<uint8_t> bytes;
  };
  int64_t next_id_ = 0;

  void append(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; out.insert(out.end(), q, q + n); }

  static FBB::Off kv(FBB& b, const std::vector<std::pair<std::string, std::string>>& md) {
    if (md.empty()) return 0;
    std::vector<FBB::Off> v;
    for (auto& e : md) {
      FBB::Off k = b.string(e.first), val = b.string(e.second);
      b.start(2);
      b.offset(0, k);
      b.offset(1, val);
      v.push_back(b.end());
    }
    return b.vec_offsets(v);
  }

  static void type_table(FBB& b, const DType& t, uint8_t* tt, FBB::Off* to) {
    switch (t.id) {
      case T_INT: b.start(2); b.scalar<int32_t>(0, t.bits, 0); b.scalar<uint8_t>(1, t.sgn ? 1 : 0, 0); *to = b.end(); *tt = FB_Int; break;
      case T_UTF8: b.start(0); *to = b.end(); *tt = FB_Utf8; break;
      case T_BINARY: b.start(0); *to = b.end(); *tt = FB_Binary; break;
      case T_UTF8VIEW: b.start(0); *to = b.end(); *tt = FB_Utf8View; break;
      case T_FSB: b.start(1); b.scalar<int32_t>(0, t.width, 0); *to = b.end(); *tt = FB_FSB; break;
      case T_TIMESTAMP_NS_UTC: {
        FBB::Off tz = b.string("UTC");
        b.start(2); b.scalar<int16_t>(0, 3 /*NANOSECOND*/, 0); b.offset(1, tz); *to = b.end(); *tt = FB_Timestamp; break;
      }
      case T_STRUCT: b.start(0); *to = b.end(); *tt = FB_Struct; break;
      case T_LISTVIEW: b.start(0); *to = b.end(); *tt = FB_ListView; break;
      case T_LIST: b.start(0); *to = b.end(); *tt = FB_List; break;
      case T_BOOL: b.start(0); *to = b.end(); *tt = FB_Bool; break;
      case T_REE: b.start(0); *to = b.end(); *tt = FB_REE; break;
      case T_DICT_U32: assert(false); break;
    }
  }

  FBB::Off field(FBB& b, const Field& f) {
    const DType* t = f.type.get();
    FBB::Off de = 0;
    if (t->id == T_DICT_U32) {
      int64_t id = next_id_++;
      b.start(2); b.scalar<int32_t>(0, 32, 0); FBB::Off it = b.end();  // Int{32, unsigned}
      b.start(4); b.scalar<int64_t>(0, id, 0); b.offset(1, it); de = b.end();
      t = t->dict_value.get();
    }
    std::vector<FBB::Off> ch;
    for (auto& c : t->kids) ch.push_back(field(b, c));
    FBB::Off chv = b.vec_offsets(ch);
    FBB::Off md = kv(b, f.metadata);
    uint8_t tt = 0; FBB::Off to = 0;
    type_table(b, *t, &tt, &to);
    FBB::Off name = b.string(f.name);
    b.start(7);
    b.offset(0, name);
    b.scalar<uint8_t>(1, f.nullable ? 1 : 0, 0);
    b.scalar<uint8_t>(2, tt, 0);
    b.offset(3, to);
    b.offset(4, de);
    b.offset(5, chv);
    b.offset(6, md);
    return b.end();
  }

  static void message(FBB& b, uint8_t header_type, FBB::Off header, int64_t body_len) {
    b.start(5);
    b.scalar<int16_t>(0, 4 /*MetadataVersion V5*/, 0);
    b.scalar<uint8_t>(1, header_type, 0);
    b.offset(2, header);
    b.scalar<int64_t>(3, body_len, 0);
    FBB::Off m = b.end();
    b.finish(m);
  }

  void emit(const FBB& b, const uint8_t* body, size_t body_len) {
    uint32_t n = b.size();
    uint32_t padded = ((n + 8 + 7) & ~7u) - 8;
    uint32_t pre[2] = {0xFFFFFFFFu, padded};
    append(pre, 8);
    append(b.data(), n);
    static const uint8_t z[8] = {0};
    append(z, padded - n);
    if (body_len) append(body, body_len);
  }

  static void put(Body& body, const Buf& buf) {
    int64_t off = (int64_t)body.bytes.size();
    int64_t n = buf ? (int64_t)buf->size() : 0;
    body.buffers.push_back(off);
    body.buffers.push_back(n);
    if (n) body.bytes.insert(body.bytes.end(), buf->begin(), buf->end());
    body.bytes.resize((body.bytes.size() + 7) & ~(size_t)7, 0);
  }

  // depth-first encoding of one array: its FieldNode, its buffers, then its children
  static void encode(const ArrayData& a, Body& body) {
    body.nodes.push_back(a.len);
    body.nodes.push_back(a.nulls);
    TypeId id = a.type->id;
    if (id != T_REE) put(body, a.nulls ? a.bufs[0] : nullptr);
    switch (id) {
      case T_INT: case T_FSB: case T_TIMESTAMP_NS_UTC: case T_DICT_U32: case T_BOOL: case T_LIST:
        put(body, a.bufs[1]);
        break;
      case T_UTF8: case T_BINARY: case T_LISTVIEW:
        put(body, a.bufs[1]);
        put(body, a.bufs[2]);
        break;
      case T_UTF8VIEW:
        put(body, a.bufs[1]);
        for (size_t i = 2; i < a.bufs.size(); i++) put(body, a.bufs[i]);
        body.variadic.push_back((int64_t)a.bufs.size() - 2);
        break;
      case T_STRUCT: case T_REE:
        break;
    }
    for (auto& k : a.kids) encode(k, body);
  }

  static FBB::Off record_batch(FBB& b, int64_t length, const Body& body) {
    FBB::Off var = body.variadic.empty() ? 0 : b.vec_i64(body.variadic, 1);
    FBB::Off bufs = b.vec_i64(body.buffers, 2);
    FBB::Off nodes = b.vec_i64(body.nodes, 2);
    b.start(5);
    b.scalar<int64_t>(0, length, 0);
    b.offset(1, nodes);
    b.offset(2, bufs);
    b.offset(4, var);
    return b.end();
  }

  // walk (type, array) in schema order: a dictionary's nested dictionaries first, then itself
  void dicts(const TypeP& t, const ArrayData& a, int64_t& id) {
    const DType* vt = t.get();
    const ArrayData* va = &a;
    int64_t my = -1;
    if (t->id == T_DICT_U32) {
      my = id++;
      vt = t->dict_value.get();
      va = a.dict.get();
    }
    for (size_t i = 0; i < vt->kids.size(); i++) dicts(vt->kids[i].type, va->kids[i], id);
    if (my >= 0) {
      Body body;
      encode(*a.dict, body);
      FBB b;
      FBB::Off rb = record_batch(b, a.dict->len, body);
      b.start(3);
      b.scalar<int64_t>(0, my, 0);
      b.offset(1, rb);
      FBB::Off db = b.end();
      message(b, FB_DictionaryBatch, db, (int64_t)body.bytes.size());
      emit(b, body.bytes.data(), body.bytes.size());
    }
  }
};

}  // namespace orc
#endif
