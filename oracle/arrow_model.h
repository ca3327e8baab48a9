/*
 * oracle/arrow_model.h — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A small in-memory Arrow array model plus builders that restate the *observable* semantics of
 * the github.com/apache/arrow-go/v18 v18.5.2 builders the reference uses (go.mod:12). arrow-go
 * is not vendored under /root/reference, so each rule below is a restatement of the published
 * Arrow columnar format + the arrow-go builder behaviour at the reference's call sites
 * (reporter/arrow.go, reporter/arrow_v2.go); byte-level agreement with a Go build is UNPINNED
 * (see DESIGN.md "parity unpinned"). Logical agreement is pinned by decoding with pyarrow.
 */
#ifndef ORACLE_ARROW_MODEL_H
#define ORACLE_ARROW_MODEL_H
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

namespace orc {

enum TypeId { T_INT, T_UTF8, T_UTF8VIEW, T_FSB, T_TIMESTAMP_NS_UTC, T_STRUCT, T_LISTVIEW, T_REE, T_DICT_U32, T_BINARY, T_LIST, T_BOOL };

struct DType;
using TypeP = std::shared_ptr<DType>;
struct Field {
  std::string name;
  TypeP type;
  bool nullable = false;
  std::vector<std::pair<std::string, std::string>> metadata;
};
struct DType {
  TypeId id;
  int bits = 0;      // T_INT
  bool sgn = false;  // T_INT
  int width = 0;     // T_FSB
  std::vector<Field> kids;  // struct fields / listview item / ree {run_ends, values}
  TypeP dict_value;         // T_DICT_U32: value type (index type is always uint32 on this path)
};
inline TypeP mk(TypeId id) { auto t = std::make_shared<DType>(); t->id = id; return t; }
inline TypeP int_t(int bits, bool sgn) { auto t = mk(T_INT); t->bits = bits; t->sgn = sgn; return t; }
inline TypeP dict_t(TypeP v) { auto t = mk(T_DICT_U32); t->dict_value = std::move(v); return t; }
inline TypeP struct_t(std::vector<Field> f) { auto t = mk(T_STRUCT); t->kids = std::move(f); return t; }
// arrow.ListViewOf(t): element field "item", nullable
inline TypeP listview_t(TypeP v) { auto t = mk(T_LISTVIEW); t->kids = {Field{"item", std::move(v), true, {}}}; return t; }
// arrow.ListOf(t): element field "item", nullable
inline TypeP list_t(TypeP v) { auto t = mk(T_LIST); t->kids = {Field{"item", std::move(v), true, {}}}; return t; }
// arrow.RunEndEncodedOf(runEnds, values): children "run_ends" (non-null) and "values" (nullable)
inline TypeP ree_t(TypeP v) {
  auto t = mk(T_REE);
  t->kids = {Field{"run_ends", int_t(32, true), false, {}}, Field{"values", std::move(v), true, {}}};
  return t;
}

using Buf = std::shared_ptr<std::vector<uint8_t>>;
inline Buf mkbuf(size_t n = 0) { return std::make_shared<std::vector<uint8_t>>(n); }
template <class T>
Buf buf_of(const std::vector<T>& v) {
  auto b = mkbuf(v.size() * sizeof(T));
  if (!v.empty()) memcpy(b->data(), v.data(), v.size() * sizeof(T));
  return b;
}

// bufs[0] is always the validity bitmap slot (nullptr when the array has no nulls — arrow-go
// leaves it nil and the IPC writer emits a zero-length buffer for it).
struct ArrayData {
  TypeP type;
  int64_t len = 0;
  int64_t nulls = 0;
  std::vector<Buf> bufs;
  std::vector<ArrayData> kids;
  std::shared_ptr<ArrayData> dict;  // T_DICT_U32: dictionary values
};

inline Buf pack_validity(const std::vector<uint8_t>& valid, int64_t nulls) {
  if (nulls == 0) return nullptr;
  auto b = mkbuf((valid.size() + 7) / 8);
  for (size_t i = 0; i < valid.size(); i++)
    if (valid[i]) (*b)[i >> 3] |= (uint8_t)(1u << (i & 7));
  return b;
}

// array.{Int32,Uint32,Int64,Uint64,Timestamp}Builder: Append / AppendNull (slot left zero) / NewArray
template <class T>
struct PrimBuilder {
  std::vector<T> v;
  std::vector<uint8_t> valid;
  int64_t nulls = 0;
  void Append(T x) { v.push_back(x); valid.push_back(1); }
  void AppendNull() { v.push_back(T()); valid.push_back(0); nulls++; }
  int Len() const { return (int)v.size(); }
  bool IsNull(int i) const { return !valid[i]; }
  T Value(int i) const { return v[i]; }
  ArrayData NewArray(TypeP t) {
    ArrayData a;
    a.type = std::move(t);
    a.len = (int64_t)v.size();
    a.nulls = nulls;
    a.bufs = {pack_validity(valid, nulls), buf_of(v)};
    v.clear(); valid.clear(); nulls = 0;
    return a;
  }
};

// array.BooleanBuilder (no nulls on this path): one bit per value
struct BoolBuilder {
  std::vector<uint8_t> v;
  void Append(bool x) { v.push_back(x ? 1 : 0); }
  ArrayData NewArray() {
    ArrayData a;
    a.type = mk(T_BOOL);
    a.len = (int64_t)v.size();
    a.bufs = {nullptr, pack_validity(v, 1)};
    if (!a.bufs[1]) a.bufs[1] = mkbuf(0);
    v.clear();
    return a;
  }
};

// array.ListBuilder: Append(valid) records the child length as the entry's start offset; the closing offset
// is appended when the array is built (n+1 int32 offsets). The child array is supplied by the caller.
struct ListBuilder {
  std::vector<int32_t> off;
  std::vector<uint8_t> valid;
  int64_t nulls = 0;
  void Append(bool v, int64_t child_len) { off.push_back((int32_t)child_len); valid.push_back(v ? 1 : 0); if (!v) nulls++; }
  ArrayData NewArray(TypeP type, ArrayData child) {
    off.push_back((int32_t)child.len);
    ArrayData a;
    a.type = std::move(type);
    a.len = (int64_t)valid.size();
    a.nulls = nulls;
    a.bufs = {pack_validity(valid, nulls), buf_of(off)};
    a.kids = {std::move(child)};
    off.clear(); valid.clear(); nulls = 0;
    return a;
  }
};

// array.StringBuilder (utf8, int32 offsets)
struct StringBuilder {
  std::vector<int32_t> off{0};
  std::vector<uint8_t> data;
  std::vector<uint8_t> valid;
  int64_t nulls = 0;
  void Append(std::string_view s) {
    data.insert(data.end(), s.begin(), s.end());
    off.push_back((int32_t)data.size());
    valid.push_back(1);
  }
  void AppendNull() { off.push_back((int32_t)data.size()); valid.push_back(0); nulls++; }
  int Len() const { return (int)valid.size(); }
  bool IsNull(int i) const { return !valid[i]; }
  std::string_view Value(int i) const {
    return std::string_view((const char*)data.data() + off[i], (size_t)(off[i + 1] - off[i]));
  }
  ArrayData NewArray(TypeId id = T_UTF8) {  // T_UTF8 (string) or T_BINARY: identical layout, different type tag
    ArrayData a;
    a.type = mk(id);
    a.len = (int64_t)valid.size();
    a.nulls = nulls;
    a.bufs = {pack_validity(valid, nulls), buf_of(off), buf_of(data)};
    off.assign(1, 0); data.clear(); valid.clear(); nulls = 0;
    return a;
  }
};

// array.BinaryDictionaryBuilder over Dictionary<uint32, utf8>: memo table assigns dictionary
// indices in first-insertion order; AppendNull appends a null *index* and no dictionary entry.
struct BinaryDictBuilder {
  std::unordered_map<std::string, uint32_t> memo;
  std::vector<std::string> values;
  PrimBuilder<uint32_t> idx;
  void Append(std::string_view s) {
    auto it = memo.find(std::string(s));
    uint32_t i;
    if (it == memo.end()) {
      i = (uint32_t)values.size();
      memo.emplace(std::string(s), i);
      values.emplace_back(s);
    } else {
      i = it->second;
    }
    idx.Append(i);
  }
  void AppendNull() { idx.AppendNull(); }
  int Len() const { return idx.Len(); }
  const std::string& Value(uint32_t i) const { return values[i]; }
  ArrayData NewArray(TypeId value_id = T_UTF8) {
    StringBuilder sb;
    for (auto& s : values) sb.Append(s);
    auto d = std::make_shared<ArrayData>(sb.NewArray(value_id));
    ArrayData a = idx.NewArray(dict_t(mk(value_id)));
    a.dict = d;
    memo.clear(); values.clear();
    return a;
  }
};

// array.StringViewBuilder: 16-byte views; strings <= 12 bytes inline, longer ones go to
// 32 KiB data blocks (first block with room wins, else a new block) — arrow-go
// array/bufferbuilder "multiBufferBuilder" (UNPINNED dependency behaviour).
struct StringViewBuilder {
  static constexpr size_t kBlock = 32 << 10;
  struct Block { std::vector<uint8_t> d; size_t cap; };
  std::vector<uint8_t> views;
  std::vector<Block> blocks;
  size_t cur = 0;
  std::vector<uint8_t> valid;
  int64_t nulls = 0;
  static size_t up64(size_t n) { return (n + 63) & ~(size_t)63; }
  void reserve(size_t n) {
    if (blocks.empty()) { blocks.push_back(Block{{}, up64(n < kBlock ? kBlock : n)}); cur = 0; return; }
    if (n <= blocks[cur].cap - blocks[cur].d.size()) return;
    for (size_t i = 0; i < blocks.size(); i++)
      if (n <= blocks[i].cap - blocks[i].d.size()) { cur = i; return; }
    blocks.push_back(Block{{}, up64(n < kBlock ? kBlock : n)});
    cur = blocks.size() - 1;
  }
  void Append(std::string_view s) {
    uint8_t v[16] = {0};
    int32_t n = (int32_t)s.size();
    memcpy(v, &n, 4);
    if (s.size() <= 12) {
      memcpy(v + 4, s.data(), s.size());
    } else {
      reserve(s.size());
      int32_t bi = (int32_t)cur, of = (int32_t)blocks[cur].d.size();
      memcpy(v + 4, s.data(), 4);
      memcpy(v + 8, &bi, 4);
      memcpy(v + 12, &of, 4);
      blocks[cur].d.insert(blocks[cur].d.end(), s.begin(), s.end());
    }
    views.insert(views.end(), v, v + 16);
    valid.push_back(1);
  }
  void AppendNull() { views.insert(views.end(), 16, 0); valid.push_back(0); nulls++; }
  ArrayData NewArray() {
    ArrayData a;
    a.type = mk(T_UTF8VIEW);
    a.len = (int64_t)valid.size();
    a.nulls = nulls;
    a.bufs = {pack_validity(valid, nulls), buf_of(views)};
    for (auto& b : blocks) a.bufs.push_back(buf_of(b.d));
    views.clear(); blocks.clear(); valid.clear(); nulls = 0; cur = 0;
    return a;
  }
};

// array.RunEndEncodedBuilder run bookkeeping: run ends are emitted lazily by finishRun().
struct ReeCore {
  int64_t length = 0;
  std::vector<int32_t> run_ends;
  void finishRun() { if (length == 0) return; run_ends.push_back((int32_t)length); }
  void Append(uint64_t n) { finishRun(); length += (int64_t)n; }
  void ContinueRun(uint64_t n) { length += (int64_t)n; }
  int64_t Len() const { return length; }
  ArrayData wrap(TypeP value_type, ArrayData values) {
    ArrayData re;
    re.type = int_t(32, true);
    re.len = (int64_t)run_ends.size();
    re.bufs = {nullptr, buf_of(run_ends)};
    ArrayData a;
    a.type = ree_t(std::move(value_type));
    a.len = length;
    a.kids = {std::move(re), std::move(values)};
    run_ends.clear();
    length = 0;
    return a;
  }
};

}  // namespace orc
#endif
