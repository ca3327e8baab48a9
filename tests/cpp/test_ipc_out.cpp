// CPU-tier test of the PRODUCT's IPC stream planner (parca_agent_b200/csrc/ipc_out.hpp) with host-resident buffers: the
// same toy record is built once as pa::Node trees over BufRef::host / BufRef::zeros and once with the oracle's array
// model + writer (oracle/arrow_model.h, oracle/ipc_writer.h); the two streams must be byte-identical, and the Python side
// of the test decodes the result with pyarrow. Covers: flatbuffer metadata construction, dictionary ids in schema
// pre-order with a dictionary nested inside a dictionary's value struct, inner-first dictionary batches, REE parents
// without buffers, List / ListView / Struct / Bool / Binary / FixedSizeBinary / StringView(variadic) nodes, validity
// buffers only when null_count > 0, zero-filled buffers, 8-byte padding.
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../oracle/arrow_model.h"
#include "../../oracle/ipc_writer.h"
#include "../../parca_agent_b200/csrc/ipc_out.hpp"

using pa::BufRef;
using pa::Node;
using pa::Ty;

template <class T>
static BufRef href(const std::vector<T>& v) { return BufRef::host(v.data(), v.size() * sizeof(T)); }

int main(int argc, char** argv) {
  const int64_t N = 5;
  // ---- shared raw buffers -------------------------------------------------------------------------------------
  std::vector<int64_t> value = {10, -20, 30, 40, 50};
  std::vector<uint8_t> value_valid = {0x1B};                                  // rows 0,1,3,4 valid; row 2 null
  std::vector<int32_t> run_ends = {2, 5};
  std::vector<uint32_t> ree_idx = {1, 0};
  std::vector<int32_t> dict_off = {0, 3, 8};
  std::string dict_bytes = "abchello";
  std::vector<int32_t> list_off = {0, 2, 2, 3, 3, 4};
  std::vector<uint8_t> list_valid = {0x17};                                   // entry 3 null
  std::vector<uint64_t> addr = {0x1000, 0x2000, 0x3000, 0x4000};
  std::vector<uint32_t> fn_idx = {0, 1, 1, 0};                                 // dict<u32, struct{name: utf8view, file: dict<u32, utf8>}>
  std::vector<uint8_t> views(2 * 16, 0);
  const std::string long_name = "a_function_name_longer_than_twelve_bytes";
  { int32_t n0 = 4; memcpy(views.data(), &n0, 4); memcpy(views.data() + 4, "main", 4);
    int32_t n1 = (int32_t)long_name.size(), bi = 0, of = 0; memcpy(views.data() + 16, &n1, 4); memcpy(views.data() + 20, long_name.data(), 4);
    memcpy(views.data() + 24, &bi, 4); memcpy(views.data() + 28, &of, 4); }
  std::vector<uint32_t> file_idx = {0, 0};
  std::vector<int32_t> file_off = {0, 4};
  std::string file_bytes = "a.py";
  std::vector<uint8_t> flags = {0x15};                                         // bool column: T F T F T
  std::vector<uint8_t> ids(N * 16);
  for (size_t i = 0; i < ids.size(); i++) ids[i] = (uint8_t)(i * 7);
  std::vector<int32_t> lv_off = {0, 1, 1, 0, 3}, lv_size = {1, 2, 0, 3, 1};

  // ---- product: Node trees ---------------------------------------------------------------------------------------
  auto int_node = [](const char* name, int bits, bool sgn, bool nullable, int64_t len, BufRef data) {
    Node n; n.ty = Ty::Int; n.name = name; n.bits = bits; n.is_signed = sgn; n.nullable = nullable; n.length = len; n.bufs = {data}; return n; };
  auto str_node = [](const char* name, Ty ty, int64_t len, BufRef off, BufRef data) {
    Node n; n.ty = ty; n.name = name; n.nullable = true; n.length = len; n.bufs = {off, data}; return n; };
  std::vector<Node> cols;
  { Node v = int_node("value", 64, true, true, N, href(value)); v.null_count = 1; v.validity = href(value_valid); cols.push_back(std::move(v)); }
  { Node dictv = str_node("values", Ty::Binary, 2, href(dict_off), BufRef::host(dict_bytes.data(), dict_bytes.size()));
    Node values; values.ty = Ty::DictU32; values.name = "values"; values.nullable = true; values.length = 2; values.bufs = {href(ree_idx)};
    values.dict.reset(new Node(std::move(dictv)));
    Node ree; ree.ty = Ty::RunEnd; ree.name = "label"; ree.nullable = true; ree.length = N;
    ree.kids.push_back(int_node("run_ends", 32, true, false, 2, href(run_ends)));
    ree.kids.push_back(std::move(values));
    cols.push_back(std::move(ree)); }
  { Node name; name.ty = Ty::Utf8View; name.name = "name"; name.nullable = true; name.length = 2;
    name.bufs = {BufRef::host(views.data(), views.size()), BufRef::host(long_name.data(), long_name.size())};
    Node filev = str_node("file", Ty::Utf8, 1, href(file_off), BufRef::host(file_bytes.data(), file_bytes.size()));
    Node file; file.ty = Ty::DictU32; file.name = "file"; file.nullable = true; file.length = 2; file.bufs = {href(file_idx)}; file.dict.reset(new Node(std::move(filev)));
    Node fstruct; fstruct.ty = Ty::Struct; fstruct.name = "function"; fstruct.length = 2;
    fstruct.kids.push_back(std::move(name)); fstruct.kids.push_back(std::move(file));
    fstruct.kids.push_back(int_node("start_line", 64, false, false, 2, BufRef::zeros(16)));
    Node fdict; fdict.ty = Ty::DictU32; fdict.name = "function"; fdict.length = 4; fdict.bufs = {href(fn_idx)}; fdict.dict.reset(new Node(std::move(fstruct)));
    Node item; item.ty = Ty::Struct; item.name = "item"; item.nullable = true; item.length = 4;
    item.kids.push_back(int_node("address", 64, false, false, 4, href(addr)));
    item.kids.push_back(std::move(fdict));
    Node lst; lst.ty = Ty::List; lst.name = "locations"; lst.nullable = false; lst.length = N; lst.null_count = 1; lst.validity = href(list_valid);
    lst.bufs = {href(list_off)}; lst.kids.push_back(std::move(item));
    cols.push_back(std::move(lst)); }
  { Node b; b.ty = Ty::Bool; b.name = "is_complete"; b.length = N; b.bufs = {href(flags)}; cols.push_back(std::move(b)); }
  { Node id; id.ty = Ty::FixedBinary; id.name = "stacktrace_id"; id.byte_width = 16; id.length = N; id.bufs = {href(ids)};
    id.metadata = {{"ARROW:extension:name", "arrow.uuid"}, {"ARROW:extension:metadata", ""}}; cols.push_back(std::move(id)); }
  { Node item = int_node("item", 32, false, true, 4, href(fn_idx));
    Node lv; lv.ty = Ty::ListView; lv.name = "view"; lv.nullable = true; lv.length = N; lv.bufs = {href(lv_off), href(lv_size)}; lv.kids.push_back(std::move(item));
    cols.push_back(std::move(lv)); }
  pa::StreamPlan plan;
  plan.build(cols, {{"parca_write_schema_version", "test"}}, N);
  std::vector<uint8_t> got(plan.total, 0xCD);  // poison: every byte must be written (padding included)
  for (auto& p : plan.placements) if (p.src.kind == BufRef::DEVICE) { fprintf(stderr, "unexpected device buffer\n"); return 2; }
  plan.write_host_parts(got.data());
  // StreamPlan::tail_distances (what the early copy-out of a flush relies on): for every suffix of the column list, the
  // end-relative position it predicts for each buffer must be where the full plan put that buffer
  size_t checked = 0;
  for (size_t first = 0; first < cols.size(); first++) {
    for (auto& bd : pa::StreamPlan::tail_distances(&cols[first], cols.size() - first)) {
      if (!bd.first.len) continue;
      bool found = false;
      for (auto& p : plan.placements) if (p.src.ptr == bd.first.ptr && p.src.len == bd.first.len && plan.total - p.at == bd.second) found = true;
      if (!found) { fprintf(stderr, "tail_distances: a buffer of the columns from %zu on is not where the plan has it\n", first); return 3; }
      checked++;
    }
  }
  if (checked < cols.size()) { fprintf(stderr, "tail_distances: nothing was checked\n"); return 3; }

  // ---- oracle: the same record with the checker's model --------------------------------------------------------------
  using namespace orc;
  auto raw = [](const void* p, size_t n) { auto b = mkbuf(n); if (n) memcpy(b->data(), p, n); return b; };
  auto prim = [&](TypeP t, int64_t len, Buf data) { ArrayData a; a.type = std::move(t); a.len = len; a.bufs = {nullptr, std::move(data)}; return a; };
  auto strs = [&](TypeId id, int64_t len, const std::vector<int32_t>& off, const std::string& bytes) {
    ArrayData a; a.type = mk(id); a.len = len; a.bufs = {nullptr, buf_of(off), raw(bytes.data(), bytes.size())}; return a; };
  std::vector<Field> fields;
  std::vector<ArrayData> ocols;
  { ArrayData v = prim(int_t(64, true), N, buf_of(value)); v.nulls = 1; v.bufs[0] = buf_of(value_valid); fields.push_back(Field{"value", v.type, true, {}}); ocols.push_back(std::move(v)); }
  { ArrayData values = prim(dict_t(mk(T_BINARY)), 2, buf_of(ree_idx)); values.dict = std::make_shared<ArrayData>(strs(T_BINARY, 2, dict_off, dict_bytes));
    ArrayData re = prim(int_t(32, true), 2, buf_of(run_ends));
    ArrayData ree; ree.type = ree_t(dict_t(mk(T_BINARY))); ree.len = N; ree.kids = {std::move(re), std::move(values)};
    fields.push_back(Field{"label", ree.type, true, {}}); ocols.push_back(std::move(ree)); }
  { TypeP fn_t = struct_t({Field{"name", mk(T_UTF8VIEW), true, {}}, Field{"file", dict_t(mk(T_UTF8)), true, {}}, Field{"start_line", int_t(64, false), false, {}}});
    TypeP item_t = struct_t({Field{"address", int_t(64, false), false, {}}, Field{"function", dict_t(fn_t), false, {}}});
    ArrayData name; name.type = mk(T_UTF8VIEW); name.len = 2; name.bufs = {nullptr, buf_of(views), raw(long_name.data(), long_name.size())};
    ArrayData file = prim(dict_t(mk(T_UTF8)), 2, buf_of(file_idx)); file.dict = std::make_shared<ArrayData>(strs(T_UTF8, 1, file_off, file_bytes));
    ArrayData fstruct; fstruct.type = fn_t; fstruct.len = 2; fstruct.bufs = {nullptr};
    fstruct.kids = {std::move(name), std::move(file), prim(int_t(64, false), 2, mkbuf(16))};
    ArrayData fdict = prim(dict_t(fn_t), 4, buf_of(fn_idx)); fdict.dict = std::make_shared<ArrayData>(std::move(fstruct));
    ArrayData item; item.type = item_t; item.len = 4; item.bufs = {nullptr}; item.kids = {prim(int_t(64, false), 4, buf_of(addr)), std::move(fdict)};
    ArrayData lst; lst.type = list_t(item_t); lst.len = N; lst.nulls = 1; lst.bufs = {buf_of(list_valid), buf_of(list_off)}; lst.kids = {std::move(item)};
    fields.push_back(Field{"locations", lst.type, false, {}}); ocols.push_back(std::move(lst)); }
  { ArrayData b; b.type = mk(T_BOOL); b.len = N; b.bufs = {nullptr, buf_of(flags)}; fields.push_back(Field{"is_complete", b.type, false, {}}); ocols.push_back(std::move(b)); }
  { ArrayData id; id.type = mk(T_FSB); id.type->width = 16; id.len = N; id.bufs = {nullptr, buf_of(ids)};
    fields.push_back(Field{"stacktrace_id", id.type, false, {{"ARROW:extension:name", "arrow.uuid"}, {"ARROW:extension:metadata", ""}}}); ocols.push_back(std::move(id)); }
  { TypeP lvt = listview_t(int_t(32, false));
    ArrayData lv; lv.type = lvt; lv.len = N; lv.bufs = {nullptr, buf_of(lv_off), buf_of(lv_size)}; lv.kids = {prim(int_t(32, false), 4, buf_of(fn_idx))};
    fields.push_back(Field{"view", lvt, true, {}}); ocols.push_back(std::move(lv)); }
  IpcWriter iw;
  iw.write_stream(fields, {{"parca_write_schema_version", "test"}}, ocols, N);

  if (iw.out.size() != got.size() || memcmp(iw.out.data(), got.data(), got.size()) != 0) {
    size_t i = 0;
    while (i < got.size() && i < iw.out.size() && got[i] == iw.out[i]) i++;
    fprintf(stderr, "streams differ: product %zu bytes, oracle %zu bytes, first difference at %zu\n", got.size(), iw.out.size(), i);
    return 1;
  }
  if (argc > 1) std::ofstream(argv[1], std::ios::binary).write((const char*)got.data(), (std::streamsize)got.size());
  printf("ok %zu\n", got.size());
  return 0;
}
