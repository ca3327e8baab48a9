/*
 * oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Single-threaded CPU restatement of parca-agent's V2 sample → Arrow path, used (a) as the
 * parity checker for the CUDA path and (b) as bench.py's cpu_baseline / --impl reference arm
 * (the Go toolchain is absent here and on the GPU box, so the reference itself cannot run).
 * Only tests/, __graft_entry__.smoke() and bench.py may load this library; the product
 * (parca_agent_b200/) never does.
 *
 * PARITY STATUS: "parity unpinned" at the byte level — the reference has no golden Arrow/IPC
 * vectors and cannot be executed. What IS pinned: the structural known answers of
 * reporter/arrow_v2_test.go and reporter/parca_reporter_test.go (tests/test_oracle_kat.py), the
 * XXH64 known answers, and the logical content via an independent pyarrow decode
 * (tests/test_oracle_pyarrow.py) against a literal Python transcription (tests/pyref.py).
 *
 * Each section cites the reference lines it follows (paths relative to parca-dev/parca-agent).
 * The same per-sample data-structure work as the Go code is kept on purpose (string-keyed label
 * builder map, per-sample label patching with integer formatting, hash maps per key class):
 * this is the timed baseline, not an optimised CPU implementation.
 */
#include <algorithm>
#include <cstdio>
#include <map>
#include <string>
#include <list>
#include <unordered_map>
#include <vector>

#include "../include/parcaagg.h"
#include "arrow_model.h"
#include "ipc_writer.h"
#include "xxh64.h"

namespace orc {

// ------------------------------------------------------------------------------------------
// reporter/arrow.go:23-70 — StringRunEndBuilder
struct StringRunEndBuilder {
  ReeCore ree;
  StringBuilder sb;
  void AppendNull() { ree.finishRun(); sb.AppendNull(); ree.length += 1; }  // ree.AppendNull()
  void AppendString(std::string_view v) {                                    // :50-59
    if (sb.Len() > 0 && !sb.IsNull(sb.Len() - 1) && v == sb.Value(sb.Len() - 1)) { ree.ContinueRun(1); return; }
    ree.Append(1);
    sb.Append(v);
  }
  ArrayData NewArray() { ree.finishRun(); return ree.wrap(mk(T_UTF8), sb.NewArray()); }
};

// reporter/arrow.go:83-139 — BinaryDictionaryRunEndBuilder (label columns)
struct BinaryDictionaryRunEndBuilder {
  ReeCore ree;
  BinaryDictBuilder bd;
  int64_t Len() const { return ree.Len(); }
  void AppendNull() { ree.finishRun(); bd.AppendNull(); ree.length += 1; }  // :129-131
  void EnsureLength(int64_t l) { while (ree.Len() < l) AppendNull(); }       // :123-127
  void Append(std::string_view v) {                                          // :97-106
    int n = bd.idx.Len();
    if (n > 0 && !bd.idx.IsNull(n - 1) && v == bd.Value(bd.idx.Value(n - 1))) { ree.ContinueRun(1); return; }
    ree.Append(1);
    bd.Append(v);
  }
  ArrayData NewArray(TypeId value_id = T_UTF8) { ree.finishRun(); return ree.wrap(dict_t(mk(value_id)), bd.NewArray(value_id)); }
};

// reporter/arrow.go:153-177 / :179-207 — Uint64RunEndBuilder / Int64RunEndBuilder
template <class T>
struct IntRunEndBuilder {
  ReeCore ree;
  PrimBuilder<T> vb;
  void Append(T v) { AppendN(v, 1); }
  void AppendN(T v, uint64_t n) {  // arrow.go:170-177
    if (vb.Len() > 0 && v == vb.Value(vb.Len() - 1)) { ree.ContinueRun(n); return; }
    ree.Append(n);
    vb.Append(v);
  }
  ArrayData NewArray(bool sgn) { ree.finishRun(); return ree.wrap(int_t(64, sgn), vb.NewArray(int_t(64, sgn))); }
};

// ------------------------------------------------------------------------------------------
// reporter/arrow_v2.go:35-160 — schema types
static TypeP FilenameDictTypeV2() { return dict_t(mk(T_UTF8)); }
static TypeP FunctionFieldTypeV2() {
  return struct_t({Field{"system_name", mk(T_UTF8VIEW), true, {}}, Field{"filename", FilenameDictTypeV2(), true, {}},
                   Field{"start_line", int_t(64, false), false, {}}});
}
static TypeP LineFieldTypeV2() {
  return struct_t({Field{"line", int_t(64, false), false, {}}, Field{"column", int_t(64, false), false, {}},
                   Field{"function", dict_t(FunctionFieldTypeV2()), false, {}}});
}
static TypeP LocationTypeV2() {
  return struct_t({Field{"address", int_t(64, false), false, {}}, Field{"frame_type", dict_t(mk(T_UTF8)), true, {}},
                   Field{"mapping_file", dict_t(mk(T_UTF8)), true, {}}, Field{"mapping_build_id", dict_t(mk(T_UTF8)), true, {}},
                   Field{"lines", listview_t(LineFieldTypeV2()), true, {}}});
}
static TypeP StacktraceTypeV2() { return listview_t(dict_t(LocationTypeV2())); }

// reporter/arrow_v2.go:21-25
struct FunctionV2 {
  std::string SystemName, Filename;
  uint64_t StartLine;
  bool operator==(const FunctionV2& o) const { return SystemName == o.SystemName && Filename == o.Filename && StartLine == o.StartLine; }
};
struct FunctionV2Hash {
  size_t operator()(const FunctionV2& f) const {
    return std::hash<std::string>()(f.SystemName) * 1000003u ^ std::hash<std::string>()(f.Filename) ^ (size_t)f.StartLine;
  }
};

// reporter/arrow_v2.go:163-218 — FunctionDictBuilderV2
struct FunctionDictBuilderV2 {
  std::unordered_map<FunctionV2, uint32_t, FunctionV2Hash> index;
  StringViewBuilder sysName;
  BinaryDictBuilder filename;
  PrimBuilder<uint64_t> startLn;
  uint32_t AppendFunction(const FunctionV2& f) {  // :186-208
    auto it = index.find(f);
    if (it != index.end()) return it->second;
    uint32_t idx = (uint32_t)index.size();
    index.emplace(f, idx);
    if (f.SystemName.empty()) sysName.AppendNull(); else sysName.Append(f.SystemName);
    if (f.Filename.empty()) filename.AppendNull(); else filename.Append(f.Filename);
    startLn.Append(f.StartLine);
    return idx;
  }
  int Len() const { return (int)index.size(); }
  ArrayData NewArray() {
    ArrayData a;
    a.type = FunctionFieldTypeV2();
    a.len = (int64_t)index.size();
    a.bufs = {nullptr};
    a.kids = {sysName.NewArray(), filename.NewArray(), startLn.NewArray(int_t(64, false))};
    index.clear();
    return a;
  }
};

struct TraceHash {
  uint64_t hi, lo;
  bool operator==(const TraceHash& o) const { return hi == o.hi && lo == o.lo; }
};
struct TraceHashHasher { size_t operator()(const TraceHash& h) const { return (size_t)(h.hi ^ (h.lo * 0x9E3779B97F4A7C15ull)); } };
struct listEntryRef { int offset, listSize; };  // arrow_v2.go:28-31

// reporter/arrow_v2.go:228-497 — StacktraceDictBuilderV2
struct StacktraceDictBuilderV2 {
  std::unordered_map<TraceHash, listEntryRef, TraceHashHasher> index;
  PrimBuilder<int32_t> offsets, sizes;
  PrimBuilder<uint32_t> indices;
  PrimBuilder<uint64_t> locAddress;
  BinaryDictBuilder locFrameType, locMappingFile, locMappingID;
  PrimBuilder<int32_t> lineListOffsets;
  PrimBuilder<uint64_t> lineNumber, lineColumn;
  PrimBuilder<uint32_t> funcIndices;
  FunctionDictBuilderV2 funcDict;
  std::unordered_map<uint64_t, uint32_t> LocationIndex;  // map[libpf.Frame]uint32, keyed by the interned frame id
  int length = 0;

  template <class F>
  void AppendStacktrace(TraceHash h, const uint64_t* frames, int n, F&& appendLocation) {  // :288-322
    auto it = index.find(h);
    if (it != index.end()) {
      offsets.Append(it->second.offset);
      sizes.Append(it->second.listSize);
      length++;
      return;
    }
    int startOffset = indices.Len();
    int listSize = 0;
    for (int i = 0; i < n; i++) {
      uint32_t idx = appendLocation(frames[i]);
      indices.Append(idx);
      listSize++;
    }
    index.emplace(h, listEntryRef{startOffset, listSize});
    offsets.Append(startOffset);
    sizes.Append(listSize);
    length++;
  }
  int Len() const { return length; }
  int UniqueStacktraces() const { return (int)index.size(); }

  ArrayData NewArray() {  // :345-481
    int numLocations = locAddress.Len();
    ArrayData stOffsets = offsets.NewArray(int_t(32, true));
    ArrayData stSizes = sizes.NewArray(int_t(32, true));
    ArrayData locIndices = indices.NewArray(dict_t(LocationTypeV2()));

    ArrayData funcValues = funcDict.NewArray();
    ArrayData funcDictArr = funcIndices.NewArray(dict_t(FunctionFieldTypeV2()));
    funcDictArr.dict = std::make_shared<ArrayData>(std::move(funcValues));

    ArrayData lineNumArr = lineNumber.NewArray(int_t(64, false));
    ArrayData lineColArr = lineColumn.NewArray(int_t(64, false));
    int64_t numLines = lineNumArr.len;
    ArrayData lineStruct;
    lineStruct.type = LineFieldTypeV2();
    lineStruct.len = numLines;
    lineStruct.bufs = {nullptr};
    lineStruct.kids = {std::move(lineNumArr), std::move(lineColArr), std::move(funcDictArr)};

    std::vector<int32_t> lineOffsetsData = lineListOffsets.v;  // :386
    ArrayData lineOffsetsArr = lineListOffsets.NewArray(int_t(32, true));
    std::vector<int32_t> lineSizes((size_t)numLocations);       // :388-394
    for (int i = 0; i < numLocations; i++)
      lineSizes[i] = (i < numLocations - 1) ? lineOffsetsData[i + 1] - lineOffsetsData[i] : (int32_t)numLines - lineOffsetsData[i];
    int linesNullCount = 0;                                      // :403-418
    for (int i = 0; i < numLocations; i++) if (lineSizes[i] == 0) linesNullCount++;
    Buf linesValidity;
    if (linesNullCount > 0) {
      linesValidity = mkbuf(((size_t)numLocations + 7) / 8);
      for (int i = 0; i < numLocations; i++) if (lineSizes[i] > 0) (*linesValidity)[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    ArrayData linesList;                                         // :421-431
    linesList.type = listview_t(LineFieldTypeV2());
    linesList.len = numLocations;
    linesList.nulls = linesNullCount;
    linesList.bufs = {linesValidity, lineOffsetsArr.bufs[1], buf_of(lineSizes)};
    linesList.kids = {std::move(lineStruct)};

    ArrayData locStruct;                                         // :445-457
    locStruct.type = LocationTypeV2();
    locStruct.len = numLocations;
    locStruct.bufs = {nullptr};
    locStruct.kids = {locAddress.NewArray(int_t(64, false)), locFrameType.NewArray(), locMappingFile.NewArray(),
                      locMappingID.NewArray(), std::move(linesList)};
    locIndices.dict = std::make_shared<ArrayData>(std::move(locStruct));  // :463

    ArrayData lv;                                                // :467-480
    lv.type = StacktraceTypeV2();
    lv.len = length;
    lv.bufs = {nullptr, stOffsets.bufs[1], stSizes.bufs[1]};
    lv.kids = {std::move(locIndices)};
    index.clear(); LocationIndex.clear(); length = 0;
    return lv;
  }
};

// reporter/arrow_v2.go:500-682 — SampleWriterV2
struct SampleWriterV2 {
  std::unordered_map<std::string, BinaryDictionaryRunEndBuilder*> labelBuilders;
  StacktraceDictBuilderV2 Stacktrace;
  std::vector<uint8_t> StacktraceID;  // extensions.UUIDBuilder: 16 B per row
  PrimBuilder<int64_t> Value;
  StringRunEndBuilder Producer, SampleType, SampleUnit, PeriodType, PeriodUnit, Temporality;
  IntRunEndBuilder<int64_t> Period;
  IntRunEndBuilder<uint64_t> Duration;
  PrimBuilder<int64_t> Timestamp;
  ~SampleWriterV2() { for (auto& e : labelBuilders) delete e.second; }

  BinaryDictionaryRunEndBuilder* Label(const std::string& name) {  // :543-552
    auto it = labelBuilders.find(name);
    BinaryDictionaryRunEndBuilder* b;
    if (it == labelBuilders.end()) { b = new BinaryDictionaryRunEndBuilder(); labelBuilders.emplace(name, b); } else b = it->second;
    b->EnsureLength(Value.Len());
    return b;
  }
  void LabelAll(const std::string& name, const std::string& value) {  // :555-564
    auto it = labelBuilders.find(name);
    BinaryDictionaryRunEndBuilder* b;
    if (it == labelBuilders.end()) { b = new BinaryDictionaryRunEndBuilder(); labelBuilders.emplace(name, b); } else b = it->second;
    b->ree.Append((uint64_t)(Value.Len() - b->ree.Len()));
    b->bd.Append(value);
  }
};

// reporter/arrow.go:260-332 — SampleWriter (v1 schema: stacktrace ids only, every column run-end encoded)
struct SampleWriter {
  std::unordered_map<std::string, BinaryDictionaryRunEndBuilder*> labelBuilders;
  BinaryDictionaryRunEndBuilder StacktraceID;
  PrimBuilder<int64_t> Value;
  BinaryDictionaryRunEndBuilder Producer, SampleType, SampleUnit, PeriodType, PeriodUnit, Temporality;
  IntRunEndBuilder<int64_t> Period, Duration, Timestamp;
  ~SampleWriter() { for (auto& e : labelBuilders) delete e.second; }
  BinaryDictionaryRunEndBuilder* Label(const std::string& name) {  // arrow.go:523-532
    auto it = labelBuilders.find(name);
    BinaryDictionaryRunEndBuilder* b;
    if (it == labelBuilders.end()) { b = new BinaryDictionaryRunEndBuilder(); labelBuilders.emplace(name, b); } else b = it->second;
    b->EnsureLength(Value.Len());
    return b;
  }
  void LabelAll(const std::string& name, const std::string& value) {  // arrow.go:534-543
    auto it = labelBuilders.find(name);
    BinaryDictionaryRunEndBuilder* b;
    if (it == labelBuilders.end()) { b = new BinaryDictionaryRunEndBuilder(); labelBuilders.emplace(name, b); } else b = it->second;
    b->ree.Append((uint64_t)(Value.Len() - b->ree.Len()));
    b->bd.Append(value);
  }
};

// reporter/arrow.go:209-254, :545-589 — LocationsWriter (v1 stacktrace record) and its schema (:332-388)
static TypeP DictBinT() { return dict_t(mk(T_BINARY)); }
static TypeP LineStructT() {
  return struct_t({Field{"line", int_t(64, true), false, {}}, Field{"column", int_t(64, false), false, {}},
                   Field{"function_name", DictBinT(), false, {}}, Field{"function_system_name", DictBinT(), false, {}},
                   Field{"function_filename", ree_t(DictBinT()), false, {}}, Field{"function_start_line", int_t(64, true), false, {}}});
}
static TypeP LocationStructT() {
  return struct_t({Field{"address", int_t(64, false), false, {}}, Field{"frame_type", ree_t(DictBinT()), false, {}},
                   Field{"mapping_start", ree_t(int_t(64, false)), false, {}}, Field{"mapping_limit", ree_t(int_t(64, false)), false, {}},
                   Field{"mapping_offset", ree_t(int_t(64, false)), false, {}}, Field{"mapping_file", ree_t(DictBinT()), false, {}},
                   Field{"mapping_build_id", ree_t(DictBinT()), false, {}}, Field{"lines", list_t(LineStructT()), false, {}}});
}
struct LocationsWriter {
  BoolBuilder IsComplete;
  ListBuilder LocationsList;
  int64_t Locations = 0;  // StructBuilder: every Append(true) is one valid struct slot
  PrimBuilder<uint64_t> Address;
  BinaryDictionaryRunEndBuilder FrameType, MappingFile, MappingBuildID, FunctionFilename;
  IntRunEndBuilder<uint64_t> MappingStart, MappingLimit, MappingOffset;
  ListBuilder Lines;
  int64_t Line = 0;
  PrimBuilder<int64_t> LineNumber, FunctionStartLine;
  PrimBuilder<uint64_t> ColumnNumber;
  BinaryDictBuilder FunctionName, FunctionSystemName;

  // NewRecord (arrow.go:230-254) minus the stacktrace_id column, which the caller owns
  ArrayData NewLocations() {
    uint64_t numMappings = (uint64_t)MappingFile.Len();
    MappingStart.AppendN(0, numMappings);
    MappingLimit.AppendN(0, numMappings);
    MappingOffset.AppendN(0, numMappings);
    ArrayData line;
    line.type = LineStructT();
    line.len = Line;
    line.bufs = {nullptr};
    line.kids = {LineNumber.NewArray(int_t(64, true)), ColumnNumber.NewArray(int_t(64, false)), FunctionName.NewArray(T_BINARY),
                 FunctionSystemName.NewArray(T_BINARY), FunctionFilename.NewArray(T_BINARY), FunctionStartLine.NewArray(int_t(64, true))};
    ArrayData lines = Lines.NewArray(list_t(LineStructT()), std::move(line));
    ArrayData loc;
    loc.type = LocationStructT();
    loc.len = Locations;
    loc.bufs = {nullptr};
    loc.kids = {Address.NewArray(int_t(64, false)), FrameType.NewArray(T_BINARY), MappingStart.NewArray(false), MappingLimit.NewArray(false),
                MappingOffset.NewArray(false), MappingFile.NewArray(T_BINARY), MappingBuildID.NewArray(T_BINARY), std::move(lines)};
    return LocationsList.NewArray(list_t(LocationStructT()), std::move(loc));
  }
};

// ------------------------------------------------------------------------------------------
struct Stats { uint64_t rows, unique_stacks, locations, functions, location_indices, empty_samples; };

struct Reporter {
  pa_agg_config cfg;
  std::vector<std::pair<std::string, std::string>> externalLabels;
  std::vector<std::string> strings{std::string()};  // id 0 == ""
  std::vector<pa_frame_desc> frames;
  std::vector<std::vector<std::pair<std::string, std::string>>> labelsets;  // the labels LRU content (:569)
  std::unordered_map<TraceHash, std::pair<const uint64_t*, int>, TraceHashHasher> stacks;  // r.stacks LRU (:224-227), v2: never read
  // v1: the same LRU, persistent across intervals and read by buildStacktraceRecord: lru.SyncedLRU[libpf.TraceHash, libpf.Frames]
  // (parca_reporter.go:106, :876; github.com/elastic/go-freelru is not vendored: this is the textbook LRU its README describes —
  // Get moves the entry to the front, Add inserts at the front and evicts the entry at the back once `cap` entries are held).
  // cap = pa_agg_config.stack_cache_entries (cacheSize, main.go:630-645); 0 = unbounded.
  struct StackLru {
    uint64_t cap = 0;
    std::list<std::pair<TraceHash, std::vector<uint64_t>>> order;  // front = most recently used
    std::unordered_map<TraceHash, std::list<std::pair<TraceHash, std::vector<uint64_t>>>::iterator, TraceHashHasher> map;
    const std::vector<uint64_t>* Get(const TraceHash& h) {
      auto it = map.find(h);
      if (it == map.end()) return nullptr;
      order.splice(order.begin(), order, it->second);
      return &it->second->second;
    }
    void Add(const TraceHash& h, std::vector<uint64_t> fr) {
      if (cap && map.size() >= cap) { map.erase(order.back().first); order.pop_back(); }
      order.emplace_front(h, std::move(fr));
      map[h] = order.begin();
    }
  } known;
  SampleWriterV2* w = new SampleWriterV2();
  SampleWriter* w1 = new SampleWriter();  // v1 schema writer (r.sampleWriter)
  uint64_t emptySamples = 0;
  std::vector<uint8_t> ipc;
  Stats last{};

  const std::string& S(uint32_t sid) const { return strings[sid]; }

  // reporter/parca_reporter.go:568-632 — labelsForTID, per-sample patch of the cached labelset.
  // labels.Builder.Set(name, "") deletes; Labels() is sorted by name (prometheus v0.303.0, go.mod:29).
  std::vector<std::pair<std::string, std::string>> labelsForTID(uint32_t tid, uint32_t ls, const std::string& comm, uint32_t cpu) const {
    std::vector<std::pair<std::string, std::string>> lb = labelsets[ls];
    if ((cfg.label_flags & 7u) == 7u) return lb;  // :611-613
    auto set = [&lb](const char* name, std::string v) {
      auto it = std::find_if(lb.begin(), lb.end(), [&](const std::pair<std::string, std::string>& p) { return p.first == name; });
      if (v.empty()) { if (it != lb.end()) lb.erase(it); return; }
      if (it != lb.end()) it->second = std::move(v); else lb.emplace_back(name, std::move(v));
    };
    if (!(cfg.label_flags & PA_LABEL_DISABLE_CPU)) set("cpu", std::to_string(cpu));            // fmt.Sprint(cpu) :618
    if (!(cfg.label_flags & PA_LABEL_DISABLE_THREAD_ID)) set("thread_id", std::to_string(tid));  // :621
    if (!(cfg.label_flags & PA_LABEL_DISABLE_THREAD_COMM)) set("thread_name", comm);             // :624
    std::sort(lb.begin(), lb.end(), [](const std::pair<std::string, std::string>& a, const std::pair<std::string, std::string>& b) { return a.first < b.first; });
    return lb;
  }

  // reporter/parca_reporter.go:418-555 — appendLocationV2
  uint32_t appendLocationV2(uint64_t frame_id) {
    StacktraceDictBuilderV2& b = w->Stacktrace;
    auto it = b.LocationIndex.find(frame_id);
    if (it != b.LocationIndex.end()) return it->second;
    uint32_t idx = (uint32_t)b.LocationIndex.size();
    b.LocationIndex.emplace(frame_id, idx);
    const pa_frame_desc& f = frames[frame_id];
    b.lineListOffsets.Append(b.lineNumber.Len());
    b.locAddress.Append(f.address_or_lineno);
    bool exists = (f.flags & PA_FRAME_F_MAPPING_FILE) && (f.flags & PA_FRAME_F_EXEC_KNOWN);
    switch (f.kind) {
      case PA_FRAME_ABORT:  // :432-446
        b.locFrameType.Append(S(f.type_name_sid));
        b.locMappingFile.Append("agent-internal-error-frame");
        b.locMappingID.AppendNull();
        b.lineNumber.Append(0);
        b.lineColumn.Append(0);
        b.funcIndices.Append(b.funcDict.AppendFunction(FunctionV2{"aborted", "", 0}));
        break;
      case PA_FRAME_NATIVE:  // :449-476
        b.locFrameType.Append(S(f.type_name_sid));
        if (exists) {
          b.locMappingFile.Append(S(f.exec_file_name_sid));
          if (!S(f.exec_build_id_sid).empty()) {
            b.locMappingID.Append(S(f.exec_build_id_sid));
          } else {  // fid.StringNoQuotes(): 128-bit file id as 32 lower-case hex digits
            char hex[40];
            snprintf(hex, sizeof hex, "%016llx%016llx", (unsigned long long)f.file_id_hi, (unsigned long long)f.file_id_lo);
            b.locMappingID.Append(hex);
          }
        } else {
          b.locMappingFile.Append("UNKNOWN");
          b.locMappingID.AppendNull();
        }
        break;
      case PA_FRAME_KERNEL: {  // :478-514
        b.locFrameType.Append(S(f.type_name_sid));
        b.locMappingFile.Append("[kernel.kallsyms]");
        b.locMappingID.AppendNull();
        std::string moduleName = exists ? S(f.exec_file_name_sid) : std::string("vmlinux");
        std::string symbol;
        uint64_t lineNumber = 0;
        if (!S(f.function_name_sid).empty()) { symbol = S(f.function_name_sid); lineNumber = f.source_line; } else symbol = "UNKNOWN";
        b.lineNumber.Append(lineNumber);
        b.lineColumn.Append(0);
        b.funcIndices.Append(b.funcDict.AppendFunction(FunctionV2{symbol, moduleName, 0}));
        break;
      }
      case PA_FRAME_OOMPROF:  // :516-520 (frame type string is NativeFrame.String(), supplied as type_name_sid)
        b.locFrameType.Append(S(f.type_name_sid));
        b.locMappingFile.Append(S(f.source_file_sid));
        b.locMappingID.Append(S(f.function_name_sid));
        break;
      default: {  // :522-551 interpreted frames
        b.locFrameType.Append(S(f.type_name_sid));
        b.locMappingFile.Append(S(f.type_name_sid));
        b.locMappingID.AppendNull();
        uint64_t lineNumber = 0;
        std::string functionName, filePath;
        if (!S(f.function_name_sid).empty()) {
          functionName = S(f.function_name_sid);
          filePath = S(f.source_file_sid);
          lineNumber = f.source_line;
        } else {
          functionName = "UNREPORTED";
          filePath = "UNREPORTED";
        }
        if (filePath.empty()) filePath = "UNKNOWN";
        b.lineNumber.Append(lineNumber);
        b.lineColumn.Append(0);
        b.funcIndices.Append(b.funcDict.AppendFunction(FunctionV2{functionName, filePath, 0}));
      }
    }
    return idx;
  }

  // reporter/parca_reporter.go:368-412 — writeSampleV2
  void writeSampleV2(TraceHash hash, const uint64_t* fr, int nfr, int64_t ts,
                     const std::vector<std::pair<std::string, std::string>>& labels, int64_t value, uint64_t duration,
                     int64_t per, bool delta, const char* producer, const char* sampleType, const char* sampleUnit,
                     const char* periodType, const char* periodUnit) {
    for (auto& lbl : labels) w->Label(lbl.first)->Append(lbl.second);
    w->Stacktrace.AppendStacktrace(hash, fr, nfr, [this](uint64_t f) { return appendLocationV2(f); });
    uint8_t id[16];  // trace.Hash.Bytes(): big-endian hi‖lo
    for (int i = 0; i < 8; i++) { id[i] = (uint8_t)(hash.hi >> (56 - 8 * i)); id[8 + i] = (uint8_t)(hash.lo >> (56 - 8 * i)); }
    w->StacktraceID.insert(w->StacktraceID.end(), id, id + 16);
    w->Timestamp.Append(ts);
    w->Value.Append(value);
    w->SampleType.AppendString(sampleType);
    w->SampleUnit.AppendString(sampleUnit);
    w->PeriodType.AppendString(periodType);
    w->PeriodUnit.AppendString(periodUnit);
    w->Producer.AppendString(producer);
    w->Duration.Append(duration);
    w->Period.Append(per);
    if (delta) w->Temporality.AppendString("delta"); else w->Temporality.AppendNull();
  }

  // reporter/parca_reporter.go:219-244 + :332-366 — ReportTraceEvent / reportTraceEventV2
  void ReportTraceEvent(const pa_sample_hdr& h, const uint64_t* fr) {
    TraceHash hash{h.hash_hi, h.hash_lo};
    if (cfg.hash_mode == PA_HASH_XXH64X2) {
      hash.hi = orc_xxh64_impl(fr, (uint64_t)h.nframes * 8, 0);
      hash.lo = orc_xxh64_impl(fr, (uint64_t)h.nframes * 8, PA_XXH_SEED_LO);
    }
    if (cfg.schema == PA_SCHEMA_V1) {
      if (!known.Get(hash)) known.Add(hash, std::vector<uint64_t>(fr, fr + h.nframes));                           // :224-227
    } else if (stacks.find(hash) == stacks.end()) stacks.emplace(hash, std::make_pair(fr, (int)h.nframes));
    auto labels = labelsForTID(h.tid, h.labelset_id, S(h.comm_sid), h.cpu);                            // :229
    if (h.nframes == 0) emptySamples++;                                                                 // :237-239
    if (cfg.schema == PA_SCHEMA_V1) { reportTraceEventV1(hash, h, labels); return; }                     // :242-328
    const uint64_t second = 1000000000ull;
    const int64_t memPeriod = 512 * 1024;
    switch (h.kind) {  // :338-363
      case PA_KIND_CPU:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, 1, second, 1000000000ll / (int64_t)cfg.samples_per_second, true,
                      "parca_agent", "samples", "count", "cpu", "nanoseconds");
        break;
      case PA_KIND_OFFCPU:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, h.value, second, 0, true, "parca_agent", "wallclock", "nanoseconds", "samples", "count");
        break;
      case PA_KIND_CUDA:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, h.value, second, 1, true, "parca_agent", "cuda", "nanoseconds", "cuda", "nanoseconds");
        break;
      case PA_KIND_MEM_INUSE_OBJECTS:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, h.value, 0, memPeriod, false, "memory", "inuse_objects", "count", "space", "bytes");
        break;
      case PA_KIND_MEM_INUSE_SPACE:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, h.value, 0, memPeriod, false, "memory", "inuse_space", "bytes", "space", "bytes");
        break;
      case PA_KIND_MEM_ALLOC_OBJECTS:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, h.value, 0, memPeriod, false, "memory", "alloc_objects", "count", "space", "bytes");
        break;
      case PA_KIND_MEM_ALLOC_SPACE:
        writeSampleV2(hash, fr, h.nframes, h.timestamp_ns, labels, h.value, 0, memPeriod, false, "memory", "alloc_space", "bytes", "space", "bytes");
        break;
      default: break;
    }
  }

  // reporter/parca_reporter.go:246-328 — the v1 branch of ReportTraceEvent (writeSample closure + Origin switch)
  void reportTraceEventV1(TraceHash hash, const pa_sample_hdr& h, const std::vector<std::pair<std::string, std::string>>& labels) {
    char buf[16];  // trace.Hash.PutBytes16
    for (int i = 0; i < 8; i++) { buf[i] = (char)(hash.hi >> (56 - 8 * i)); buf[8 + i] = (char)(hash.lo >> (56 - 8 * i)); }
    auto writeSample = [&](int64_t value, int64_t duration, int64_t per, const char* producer, const char* sampleType, const char* sampleUnit,
                           const char* periodType, const char* periodUnit) {
      for (auto& lbl : labels) w1->Label(lbl.first)->Append(lbl.second);
      w1->StacktraceID.Append(std::string_view(buf, 16));
      w1->Timestamp.Append(h.timestamp_ns);
      w1->Value.Append(value);
      w1->SampleType.Append(sampleType);
      w1->SampleUnit.Append(sampleUnit);
      w1->PeriodType.Append(periodType);
      w1->PeriodUnit.Append(periodUnit);
      w1->Producer.Append(producer);
      w1->Duration.Append(duration);
      w1->Period.Append(per);
    };
    const int64_t second = 1000000000ll, persec = 1000000000ll / (int64_t)cfg.samples_per_second, memPeriod = 512 * 1024;
    switch (h.kind) {
      case PA_KIND_CPU:  // :284-287
        writeSample(1, second, persec, "parca_agent", "samples", "count", "cpu", "nanoseconds");
        w1->Temporality.Append("delta");
        break;
      case PA_KIND_OFFCPU:  // :288-291 (the v1 path keeps period = 1e9/Hz here, unlike v2)
        writeSample(h.value, second, persec, "parca_agent", "wallclock", "nanoseconds", "samples", "count");
        w1->Temporality.Append("delta");
        break;
      case PA_KIND_CUDA:  // :321-324
        writeSample(h.value, second, persec, "parca_agent", "cuda", "nanoseconds", "cuda", "nanoseconds");
        w1->Temporality.Append("delta");
        break;
      case PA_KIND_MEM_INUSE_OBJECTS: w1->Temporality.AppendNull(); writeSample(h.value, 0, memPeriod, "memory", "inuse_objects", "count", "space", "bytes"); break;
      case PA_KIND_MEM_INUSE_SPACE: w1->Temporality.AppendNull(); writeSample(h.value, 0, memPeriod, "memory", "inuse_space", "bytes", "space", "bytes"); break;
      case PA_KIND_MEM_ALLOC_OBJECTS: w1->Temporality.AppendNull(); writeSample(h.value, 0, memPeriod, "memory", "alloc_objects", "count", "space", "bytes"); break;
      case PA_KIND_MEM_ALLOC_SPACE: w1->Temporality.AppendNull(); writeSample(h.value, 0, memPeriod, "memory", "alloc_space", "bytes", "space", "bytes"); break;
      default: break;
    }
  }

  // reporter/parca_reporter.go:1545-1739 — buildStacktraceRecord, branch by branch; ids = n 16-byte big-endian hashes
  void buildStacktraceRecord(const uint8_t* ids, uint64_t n) {
    LocationsWriter w;
    const std::string unknownType = cfg.unknown_frame_type_sid ? S(cfg.unknown_frame_type_sid) : std::string("unknown");
    for (uint64_t i = 0; i < n; i++) {
      bool isComplete = true;
      TraceHash th{0, 0};  // libpf.TraceHashFromBytes
      for (int k = 0; k < 8; k++) { th.hi = (th.hi << 8) | ids[16 * i + k]; th.lo = (th.lo << 8) | ids[16 * i + 8 + k]; }
      const std::vector<uint64_t>* hit = known.Get(th);  // :1555
      if (!hit) {  // :1556-1573
        w.LocationsList.Append(true, w.Locations);
        w.Locations++;
        w.Address.Append(0);
        w.FrameType.Append(unknownType);
        w.MappingFile.AppendNull();
        w.MappingBuildID.AppendNull();
        w.Lines.Append(true, w.Line);
        w.Line++;
        w.LineNumber.Append(0);
        w.ColumnNumber.Append(0);
        w.FunctionName.Append("missing stacktrace");
        w.FunctionSystemName.Append("");
        w.FunctionFilename.AppendNull();
        w.FunctionStartLine.Append(0);
        w.IsComplete.Append(false);
        continue;
      }
      const std::vector<uint64_t>& traceInfo = *hit;
      w.LocationsList.Append(!traceInfo.empty(), w.Locations);  // :1576-1580
      for (uint64_t frame_id : traceInfo) {
        const pa_frame_desc& f = frames[frame_id];
        w.Locations++;
        w.Address.Append(f.address_or_lineno);
        auto line = [&](int64_t lineNumber, uint64_t column, std::string_view fn, const std::string* filename) {
          w.Lines.Append(true, w.Line);
          w.Line++;
          w.LineNumber.Append(lineNumber);
          w.ColumnNumber.Append(column);
          w.FunctionName.Append(fn);
          w.FunctionSystemName.Append("");
          if (filename) w.FunctionFilename.Append(*filename); else w.FunctionFilename.AppendNull();
          w.FunctionStartLine.Append(0);
        };
        bool exists = (f.flags & PA_FRAME_F_MAPPING_FILE) && (f.flags & PA_FRAME_F_EXEC_KNOWN);
        if (f.kind == PA_FRAME_ABORT) {  // :1586-1604
          w.FrameType.Append(S(f.type_name_sid));
          w.MappingFile.Append("agent-internal-error-frame");
          w.MappingBuildID.AppendNull();
          line(0, 0, "aborted", nullptr);
          continue;
        }
        switch (f.kind) {
          case PA_FRAME_NATIVE:  // :1606-1643
            w.FrameType.Append(S(f.type_name_sid));
            if (exists) {
              w.MappingFile.Append(S(f.exec_file_name_sid));
              if (!S(f.exec_build_id_sid).empty()) {
                w.MappingBuildID.Append(S(f.exec_build_id_sid));
              } else {
                char hex[40];
                snprintf(hex, sizeof hex, "%016llx%016llx", (unsigned long long)f.file_id_hi, (unsigned long long)f.file_id_lo);
                w.MappingBuildID.Append(hex);
              }
            } else {
              w.MappingFile.Append("UNKNOWN");
              w.MappingBuildID.AppendNull();
              isComplete = false;
            }
            w.Lines.Append(false, w.Line);
            break;
          case PA_FRAME_KERNEL: {  // :1644-1687
            w.FrameType.Append(S(f.type_name_sid));
            std::string moduleName = exists ? S(f.exec_file_name_sid) : std::string("vmlinux");
            std::string symbol;
            int64_t lineNumber = 0;
            if (!S(f.function_name_sid).empty()) { symbol = S(f.function_name_sid); lineNumber = (int64_t)f.source_line; }
            else { symbol = "UNKNOWN"; isComplete = false; }
            w.MappingBuildID.AppendNull();
            line(lineNumber, f.source_column, symbol, &moduleName);
            w.MappingFile.Append("[kernel.kallsyms]");
            break;
          }
          case PA_FRAME_OOMPROF:  // :1688-1694
            w.FrameType.Append(S(f.type_name_sid));
            w.MappingFile.Append(S(f.source_file_sid));
            w.MappingBuildID.Append(S(f.function_name_sid));
            w.Lines.Append(false, w.Line);
            isComplete = false;
            break;
          default: {  // :1695-1733
            w.FrameType.Append(S(f.type_name_sid));
            int64_t lineNumber = 0;
            std::string functionName, filePath;
            if (!S(f.function_name_sid).empty()) {
              functionName = S(f.function_name_sid);
              filePath = S(f.source_file_sid);
              lineNumber = (int64_t)f.source_line;
            } else {
              functionName = "UNREPORTED";
              filePath = "UNREPORTED";
              isComplete = false;
            }
            if (filePath.empty()) filePath = "UNKNOWN";
            if (!S(f.gnu_build_id_sid).empty()) {
              w.MappingFile.Append(S(f.mapping_file_name_sid));
              w.MappingBuildID.Append(S(f.gnu_build_id_sid));
            } else {
              w.MappingFile.Append(S(f.type_name_sid));
              w.MappingBuildID.AppendNull();
            }
            line(lineNumber, f.source_column, functionName, &filePath);
          }
        }
      }
      w.IsComplete.Append(isComplete);
    }
    // LocationsWriter.NewRecord (arrow.go:230-254) + IPC (:1336-1349)
    StringBuilder idb;
    for (uint64_t i = 0; i < n; i++) idb.Append(std::string_view((const char*)ids + 16 * i, 16));
    std::vector<Field> fields = {Field{"stacktrace_id", mk(T_BINARY), false, {}}, Field{"is_complete", mk(T_BOOL), false, {}},
                                 Field{"locations", list_t(LocationStructT()), false, {}}};
    std::vector<ArrayData> cols;
    cols.push_back(idb.NewArray(T_BINARY));
    cols.push_back(w.IsComplete.NewArray());
    cols.push_back(w.NewLocations());
    last = Stats{n, 0, (uint64_t)cols[2].kids[0].len, 0, 0, 0};
    IpcWriter iw;
    iw.write_stream(fields, {{"parca_write_schema_version", "v1"}}, cols, (int64_t)n);
    ipc.swap(iw.out);
  }

  // reporter/parca_reporter.go:1528-1543 (buildSampleRecord) + reporter/arrow.go:274-316 (NewRecord) + IPC :1390-1400
  void flushV1() {
    SampleWriter* cur = w1;
    w1 = new SampleWriter();
    stacks.clear();
    for (auto& l : externalLabels) cur->LabelAll(l.first, l.second);  // writeCommonLabels :1515-1519
    last = Stats{(uint64_t)cur->Value.Len(), (uint64_t)cur->StacktraceID.bd.values.size(), 0, 0, 0, emptySamples};
    emptySamples = 0;
    ipc.clear();
    if (cur->Value.Len() == 0) { delete cur; return; }  // :1387-1390
    std::vector<std::string> labelNames;
    for (auto& e : cur->labelBuilders) labelNames.push_back(e.first);
    std::sort(labelNames.begin(), labelNames.end());
    int64_t length = cur->Value.Len();
    TypeP dictBin = ree_t(dict_t(mk(T_BINARY)));
    std::vector<Field> fields;
    std::vector<ArrayData> cols;
    for (auto& name : labelNames) {
      BinaryDictionaryRunEndBuilder* b = cur->labelBuilders[name];
      b->EnsureLength(length);
      fields.push_back(Field{"labels." + name, dictBin, true, {}});  // ColumnLabelsPrefix, arrow.go:515-521
      cols.push_back(b->NewArray(T_BINARY));
    }
    auto add = [&](const char* name, TypeP t, ArrayData a) { fields.push_back(Field{name, std::move(t), false, {}}); cols.push_back(std::move(a)); };
    add("stacktrace_id", dictBin, cur->StacktraceID.NewArray(T_BINARY));  // ArrowSamplesField arrow.go:484-503
    add("value", int_t(64, true), cur->Value.NewArray(int_t(64, true)));
    add("producer", dictBin, cur->Producer.NewArray(T_BINARY));
    add("sample_type", dictBin, cur->SampleType.NewArray(T_BINARY));
    add("sample_unit", dictBin, cur->SampleUnit.NewArray(T_BINARY));
    add("period_type", dictBin, cur->PeriodType.NewArray(T_BINARY));
    add("period_unit", dictBin, cur->PeriodUnit.NewArray(T_BINARY));
    add("temporality", dictBin, cur->Temporality.NewArray(T_BINARY));
    add("period", ree_t(int_t(64, true)), cur->Period.NewArray(true));
    add("duration", ree_t(int_t(64, true)), cur->Duration.NewArray(true));
    add("timestamp", ree_t(int_t(64, true)), cur->Timestamp.NewArray(true));
    IpcWriter iw;
    iw.write_stream(fields, {{"parca_write_schema_version", "v1"}}, cols, length);
    ipc.swap(iw.out);
    delete cur;
  }

  // reporter/parca_reporter.go:1742-1764 + reporter/arrow_v2.go:612-663 + IPC :1779-1790
  void flush() {
    if (cfg.schema == PA_SCHEMA_V1) { flushV1(); return; }
    SampleWriterV2* cur = w;
    w = new SampleWriterV2();
    stacks.clear();
    for (auto& l : externalLabels) cur->LabelAll(l.first, l.second);  // writeCommonLabelsV2
    last = Stats{(uint64_t)cur->Value.Len(), (uint64_t)cur->Stacktrace.UniqueStacktraces(), (uint64_t)cur->Stacktrace.locAddress.Len(),
                 (uint64_t)cur->Stacktrace.funcDict.Len(), (uint64_t)cur->Stacktrace.indices.Len(), emptySamples};
    emptySamples = 0;
    ipc.clear();
    if (cur->Value.Len() == 0) { delete cur; return; }  // :1775-1778 skip empty batches

    // NewRecord: sorted label names, backfill, labels struct, 13 columns
    std::vector<std::string> labelNames;
    for (auto& e : cur->labelBuilders) labelNames.push_back(e.first);
    std::sort(labelNames.begin(), labelNames.end());
    int64_t length = cur->Value.Len();
    std::vector<Field> labelFields;
    ArrayData labelsArr;
    TypeP labelT = ree_t(dict_t(mk(T_UTF8)));
    for (auto& name : labelNames) {
      BinaryDictionaryRunEndBuilder* b = cur->labelBuilders[name];
      b->EnsureLength(length);
      labelFields.push_back(Field{name, labelT, true, {}});
      labelsArr.kids.push_back(b->NewArray());
    }
    labelsArr.type = struct_t(labelFields);
    labelsArr.len = length;
    labelsArr.bufs = {nullptr};

    ArrayData idArr;
    idArr.type = mk(T_FSB);
    idArr.type->width = 16;
    idArr.len = length;
    idArr.bufs = {nullptr, buf_of(cur->StacktraceID)};

    std::vector<Field> fields = {  // ArrowSamplesFieldV2 :581-604
        Field{"labels", labelsArr.type, false, {}},
        Field{"stacktrace", StacktraceTypeV2(), true, {}},
        Field{"stacktrace_id", idArr.type, false, {{"ARROW:extension:name", "arrow.uuid"}, {"ARROW:extension:metadata", ""}}},
        Field{"value", int_t(64, true), false, {}},
        Field{"producer", ree_t(mk(T_UTF8)), false, {}},
        Field{"sample_type", ree_t(mk(T_UTF8)), false, {}},
        Field{"sample_unit", ree_t(mk(T_UTF8)), false, {}},
        Field{"period_type", ree_t(mk(T_UTF8)), false, {}},
        Field{"period_unit", ree_t(mk(T_UTF8)), false, {}},
        Field{"temporality", ree_t(mk(T_UTF8)), true, {}},
        Field{"period", ree_t(int_t(64, true)), false, {}},
        Field{"duration", ree_t(int_t(64, false)), false, {}},
        Field{"timestamp", mk(T_TIMESTAMP_NS_UTC), false, {}},
    };
    std::vector<ArrayData> cols;
    cols.push_back(std::move(labelsArr));
    cols.push_back(cur->Stacktrace.NewArray());
    cols.push_back(std::move(idArr));
    cols.push_back(cur->Value.NewArray(int_t(64, true)));
    cols.push_back(cur->Producer.NewArray());
    cols.push_back(cur->SampleType.NewArray());
    cols.push_back(cur->SampleUnit.NewArray());
    cols.push_back(cur->PeriodType.NewArray());
    cols.push_back(cur->PeriodUnit.NewArray());
    cols.push_back(cur->Temporality.NewArray());
    cols.push_back(cur->Period.NewArray(true));
    cols.push_back(cur->Duration.NewArray(false));
    cols.push_back(cur->Timestamp.NewArray(mk(T_TIMESTAMP_NS_UTC)));

    IpcWriter iw;
    iw.write_stream(fields, {{"parca_write_schema_version", "v2"}}, cols, length);
    ipc.swap(iw.out);
    delete cur;
  }
};

// reporter/parca_reporter.go:190-216 — maybeFixTruncation; returns new length or -1
static bool utf8_valid(const uint8_t* s, size_t n) {  // unicode/utf8.ValidString
  size_t i = 0;
  while (i < n) {
    uint8_t c = s[i];
    if (c < 0x80) { i++; continue; }
    size_t need; uint32_t cp; uint32_t minv;
    if ((c & 0xE0) == 0xC0) { need = 1; cp = c & 0x1F; minv = 0x80; }
    else if ((c & 0xF0) == 0xE0) { need = 2; cp = c & 0x0F; minv = 0x800; }
    else if ((c & 0xF8) == 0xF0) { need = 3; cp = c & 0x07; minv = 0x10000; }
    else return false;
    if (i + need >= n) return false;  // truncated sequence
    for (size_t k = 1; k <= need; k++) {
      if ((s[i + k] & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (s[i + k] & 0x3F);
    }
    if (cp < minv || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += need + 1;
  }
  return true;
}

}  // namespace orc

using orc::Reporter;

extern "C" {

void* orc_create(const pa_agg_config* cfg) {
  if (!cfg || cfg->samples_per_second == 0) return nullptr;
  Reporter* r = new Reporter();
  r->cfg = *cfg;
  r->known.cap = cfg->stack_cache_entries;
  return r;
}
void orc_destroy(void* p) { if (!p) return; Reporter* r = (Reporter*)p; delete r->w; delete r->w1; delete r; }

int orc_register_strings(void* p, const uint8_t* bytes, const uint32_t* offsets, uint32_t n, uint32_t* first_id) {
  Reporter* r = (Reporter*)p;
  if (first_id) *first_id = (uint32_t)r->strings.size();
  for (uint32_t i = 0; i < n; i++) r->strings.emplace_back((const char*)bytes + offsets[i], offsets[i + 1] - offsets[i]);
  return 0;
}
int orc_register_frames(void* p, const pa_frame_desc* d, uint32_t n, uint64_t* first) {
  Reporter* r = (Reporter*)p;
  if (first) *first = r->frames.size();
  r->frames.insert(r->frames.end(), d, d + n);
  return 0;
}
int orc_register_labelsets(void* p, const pa_label_pair* pairs, const uint32_t* offsets, uint32_t n, uint32_t* first) {
  Reporter* r = (Reporter*)p;
  if (first) *first = (uint32_t)r->labelsets.size();
  for (uint32_t i = 0; i < n; i++) {
    std::vector<std::pair<std::string, std::string>> ls;
    for (uint32_t k = offsets[i]; k < offsets[i + 1]; k++) ls.emplace_back(r->S(pairs[k].name_sid), r->S(pairs[k].value_sid));
    r->labelsets.push_back(std::move(ls));
  }
  return 0;
}
// external labels are resolved lazily (strings may be registered after create)
int orc_set_external_labels(void* p, const pa_label_pair* pairs, uint32_t n) {
  Reporter* r = (Reporter*)p;
  r->externalLabels.clear();
  for (uint32_t i = 0; i < n; i++) r->externalLabels.emplace_back(r->S(pairs[i].name_sid), r->S(pairs[i].value_sid));
  return 0;
}
// one ReportTraceEvent call per row, in row order
int orc_ingest(void* p, const pa_sample_hdr* hdrs, const uint64_t* frames, uint64_t n) {
  Reporter* r = (Reporter*)p;
  for (uint64_t i = 0; i < n; i++) r->ReportTraceEvent(hdrs[i], frames + hdrs[i].frame_off);
  return 0;
}
int orc_flush(void* p, const uint8_t** ipc, uint64_t* len, uint64_t* stats6) {
  Reporter* r = (Reporter*)p;
  r->flush();
  *ipc = r->ipc.empty() ? nullptr : r->ipc.data();
  *len = r->ipc.size();
  if (stats6) {
    stats6[0] = r->last.rows; stats6[1] = r->last.unique_stacks; stats6[2] = r->last.locations;
    stats6[3] = r->last.functions; stats6[4] = r->last.location_indices; stats6[5] = r->last.empty_samples;
  }
  return 0;
}
// v1: buildStacktraceRecord for n 16-byte ids against the stacks seen so far
int orc_stacktraces(void* p, const uint8_t* ids, uint64_t n, const uint8_t** ipc, uint64_t* len, uint64_t* n_locations) {
  Reporter* r = (Reporter*)p;
  r->buildStacktraceRecord(ids, n);
  *ipc = r->ipc.data();
  *len = r->ipc.size();
  if (n_locations) *n_locations = r->last.locations;
  return 0;
}
uint64_t orc_xxh64(const void* data, uint64_t len, uint64_t seed) { return orc_xxh64_impl(data, len, seed); }

int64_t orc_fix_truncation(const uint8_t* s, uint64_t len, uint64_t maxLen) {
  if (orc::utf8_valid(s, len)) return (int64_t)len;
  int64_t begin = -1;
  if (len == maxLen) {
    for (uint64_t i = 0; i < 2; i++) {
      uint64_t idx = maxLen - i - 1;
      if ((s[idx] & 0xC0) != 0x80) { begin = (int64_t)idx; break; }
    }
  }
  if (begin == -1) return -1;
  if (!orc::utf8_valid(s, (size_t)begin)) return -1;
  return begin;
}

// known-answer hook for reporter/arrow_v2_test.go:13-47 (FunctionDictBuilderV2 dedup indices)
void* orc_funcdict_new() { return new orc::FunctionDictBuilderV2(); }
uint32_t orc_funcdict_append(void* p, const char* sys, const char* file, uint64_t start_line) {
  return ((orc::FunctionDictBuilderV2*)p)->AppendFunction(orc::FunctionV2{sys, file, start_line});
}
int orc_funcdict_len(void* p) { return ((orc::FunctionDictBuilderV2*)p)->Len(); }
void orc_funcdict_free(void* p) { delete (orc::FunctionDictBuilderV2*)p; }

}  // extern "C"
