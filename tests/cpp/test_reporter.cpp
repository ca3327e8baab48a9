// Host-layer tests for parca::ParcaReporter (parca_agent_b200/csrc/reporter.hpp), written to read
// like the reference's reporter tests (reporter/parca_reporter_test.go, reporter/arrow_v2_test.go).
//   ./test_reporter            — CPU only: a recording Sink checks interning / label / origin logic
//   ./test_reporter --gpu OUT  — drives the real C ABI on cuda:0 and writes the IPC stream of the
//                                arrow_v2_test.go:257-316 scenario to OUT (pytest compares it with the oracle)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../../parca_agent_b200/csrc/reporter.hpp"

using namespace parca;

static int failures = 0;
#define REQUIRE(cond) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

struct RecordingSink : Sink {
  std::vector<std::string> strings{""};
  std::vector<pa_frame_desc> frames;
  std::vector<std::vector<pa_label_pair>> labelsets;
  std::vector<pa_sample_hdr> rows;
  std::vector<std::vector<uint64_t>> row_frames;
  int flushes = 0;
  uint32_t RegisterString(const std::string& s) override { strings.push_back(s); return (uint32_t)strings.size() - 1; }
  uint64_t RegisterFrame(const pa_frame_desc& d) override { frames.push_back(d); return frames.size() - 1; }
  uint32_t RegisterLabelset(const std::vector<pa_label_pair>& p) override { labelsets.push_back(p); return (uint32_t)labelsets.size() - 1; }
  int Submit(const pa_sample_hdr& h, const uint64_t* ids) override { rows.push_back(h); row_frames.emplace_back(ids, ids + h.nframes); return 0; }
  int Flush(pa_agg_result* out) override { memset(out, 0, sizeof *out); out->n_rows = rows.size(); flushes++; rows.clear(); return 0; }
  void Release(pa_agg_result*) override {}
  std::string label(uint32_t ls, const char* name) const {
    for (auto& p : labelsets[ls]) if (strings[p.name_sid] == name) return strings[p.value_sid];
    return "";
  }
};

static Frame nativeFrame(uint64_t addr, bool mapped = false, FileID fid = {}) {
  Frame f; f.Type = FrameType{PA_FRAME_NATIVE, "native"}; f.AddressOrLineno = addr; f.MappingValid = mapped; f.MappingHasFile = mapped; f.MappingFileID = fid; return f;
}
static Frame kernelFrame(uint64_t addr, const char* fn, uint32_t line) {
  Frame f; f.Type = FrameType{PA_FRAME_KERNEL, "kernel"}; f.AddressOrLineno = addr; f.FunctionName = fn; f.SourceLine = line; return f;
}

static void testMaybeFixTruncation() {  // parca_reporter_test.go:18-41
  const std::string Chinese = "Go（又稱Golang[4]）是Google開發的一种静态强类型、編譯型、并发型，并具有垃圾回收功能的编程语言。";
  const std::string Chinese2 = "Linux是一种自由和开放源码的类Unix操作系统。";
  struct { std::string s, result; bool ok; } cases[] = {
      {"ASCII string", "ASCII string", true}, {Chinese.substr(0, 4), "", false}, {Chinese.substr(0, 48), Chinese.substr(0, 47), true},
      {Chinese2.substr(0, 48), Chinese2.substr(0, 48), true}, {Chinese2, Chinese2, true}};
  for (auto& c : cases) {
    std::string s = c.s;
    bool ok = MaybeFixTruncation(&s, 48);
    REQUIRE(ok == c.ok);
    REQUIRE(s == c.result);
  }
}

static void testLabelsAndInterning() {
  RecordingSink sink;
  Config cfg;
  cfg.nodeName = "test-node";
  cfg.labelsForPID = [](uint32_t pid, Labels* lb) {
    lb->emplace_back("comm", "proc-" + std::to_string(pid));
    lb->emplace_back("__meta_internal", "dropped");
    return pid != 666;  // relabeling drops pid 666
  };
  ParcaReporter r(&sink, cfg);
  Trace t; t.Hash = {1, 2}; t.Frames = {nativeFrame(0x1000), nativeFrame(0x2000)};
  TraceEventMeta m; m.PID = 1000; m.TID = 1234; m.CPU = 1; m.Comm = "myprocess"; m.Timestamp = 42;
  REQUIRE(r.ReportTraceEvent(&t, &m) == 0);
  m.CPU = 3;  // TestLabelsForTID_CPUCacheMismatch: the cpu travels per sample, never cached
  REQUIRE(r.ReportTraceEvent(&t, &m) == 0);
  REQUIRE(sink.rows.size() == 2 && sink.rows[0].cpu == 1 && sink.rows[1].cpu == 3 && sink.rows[0].tid == 1234);
  REQUIRE(sink.strings[sink.rows[0].comm_sid] == "myprocess");
  REQUIRE(sink.labelsets.size() == 1);  // cached per PID
  REQUIRE(sink.label(sink.rows[0].labelset_id, "node") == "test-node");
  REQUIRE(sink.label(sink.rows[0].labelset_id, "comm") == "proc-1000");
  REQUIRE(sink.label(sink.rows[0].labelset_id, "__meta_internal").empty());
  REQUIRE(sink.frames.size() == 2 && sink.row_frames[0] == sink.row_frames[1]);  // frame ids deduplicated by Frame value
  REQUIRE(sink.rows[0].kind == PA_KIND_CPU && sink.rows[0].hash_hi == 1 && sink.rows[0].hash_lo == 2);
  // custom labels: non-UTF8 key ignored, truncated value fixed, empty set falls back to the PID's labelset
  Trace t2 = t; t2.CustomLabels = {{"request", "abc"}, {"\xff\xfe", "x"}};
  REQUIRE(r.ReportTraceEvent(&t2, &m) == 0);
  REQUIRE(sink.labelsets.size() == 2 && sink.label(sink.rows[2].labelset_id, "request") == "abc");
  // relabel drop
  m.PID = 666;
  REQUIRE(r.ReportTraceEvent(&t, &m) == 0);
  REQUIRE(sink.rows.size() == 3 && r.skippedByRelabeling == 1);
  // empty stack still writes a row (:237-239)
  m.PID = 1000; Trace e; e.Hash = {9, 9};
  REQUIRE(r.ReportTraceEvent(&e, &m) == 0);
  REQUIRE(sink.rows.size() == 4 && sink.rows[3].nframes == 0 && r.emptySamples == 1);
}

static void testOrigins() {  // reportTraceEventV2 :338-363
  RecordingSink sink; Config cfg; cfg.nodeName = "n"; cfg.reportAllocs = true;
  ParcaReporter r(&sink, cfg);
  Trace t; t.Hash = {5, 5}; t.Frames = {kernelFrame(0x2000, "do_syscall_64", 100)};
  TraceEventMeta m; m.PID = 1; m.OffTime = 777;
  m.Origin = TraceOriginOffCPU; r.ReportTraceEvent(&t, &m);
  m.Origin = TraceOriginCuda; r.ReportTraceEvent(&t, &m);
  MemorySample ms{10, 4, 4096, 1024}; m.Origin = TraceOriginMemory; m.OriginData = &ms; r.ReportTraceEvent(&t, &m);
  MemorySample same{5, 5, 64, 64}; m.OriginData = &same; r.ReportTraceEvent(&t, &m);  // inuse rows suppressed when equal
  m.OriginData = nullptr; r.ReportTraceEvent(&t, &m);                                  // missing OriginData: skipped
  REQUIRE(sink.rows.size() == 2 + 4 + 2);
  REQUIRE(sink.rows[0].kind == PA_KIND_OFFCPU && sink.rows[0].value == 777 && sink.rows[1].kind == PA_KIND_CUDA);
  REQUIRE(sink.rows[2].kind == PA_KIND_MEM_INUSE_OBJECTS && sink.rows[2].value == 6);
  REQUIRE(sink.rows[3].kind == PA_KIND_MEM_INUSE_SPACE && sink.rows[3].value == 3072);
  REQUIRE(sink.rows[4].kind == PA_KIND_MEM_ALLOC_OBJECTS && sink.rows[4].value == 10);
  REQUIRE(sink.rows[5].kind == PA_KIND_MEM_ALLOC_SPACE && sink.rows[5].value == 4096);
  REQUIRE(sink.rows[6].kind == PA_KIND_MEM_ALLOC_OBJECTS && sink.rows[7].kind == PA_KIND_MEM_ALLOC_SPACE);
  REQUIRE(r.offcpuSamples == 1 && r.gpuSamples == 1 && r.memorySamples == 2);
}

static void testExecutables() {  // :449-476 with r.executables
  RecordingSink sink; Config cfg; cfg.nodeName = "n";
  ParcaReporter r(&sink, cfg);
  FileID fid{0xAA, 0xBB};
  Trace t; t.Hash = {1, 1}; t.Frames = {nativeFrame(0x1000, true, fid)};
  TraceEventMeta m; m.PID = 1;
  REQUIRE(!r.ExecutableKnown(fid));
  r.ReportTraceEvent(&t, &m);
  REQUIRE(sink.frames.size() == 1 && (sink.frames[0].flags & PA_FRAME_F_MAPPING_FILE) && !(sink.frames[0].flags & PA_FRAME_F_EXEC_KNOWN));
  r.ReportExecutable(ExecutableMetadata{fid, "/usr/bin/app", "build123"});
  REQUIRE(r.ExecutableKnown(fid));
  r.ReportTraceEvent(&t, &m);  // same Frame value, now resolvable: re-interned with the executable attached
  REQUIRE(sink.frames.size() == 2 && (sink.frames[1].flags & PA_FRAME_F_EXEC_KNOWN));
  REQUIRE(sink.strings[sink.frames[1].exec_file_name_sid] == "/usr/bin/app" && sink.strings[sink.frames[1].exec_build_id_sid] == "build123");
  r.ReportTraceEvent(&t, &m);
  REQUIRE(sink.frames.size() == 2 && sink.row_frames[1] == sink.row_frames[2]);
}

static void testOomprofSampleEvents() {  // SampleEvents, parca_reporter.go:709-758
  RecordingSink sink; Config cfg; cfg.nodeName = "n"; cfg.reportAllocs = false;
  ParcaReporter r(&sink, cfg);
  OomprofSampleMeta meta; meta.Timestamp = 99; meta.Comm = "victim"; meta.PID = 4242; meta.BuildID = "abcdef01"; meta.ExecutablePath = "/srv/bin/victim";
  meta.CustomLabels = {{"tenant", "blue"}};
  MemorySample a; a.Allocs = 10; a.Frees = 3; a.AllocBytes = 4096; a.FreeBytes = 96; a.Addresses = {0x1000, 0x2000, 0x1000};
  MemorySample b; b.Allocs = 2; b.Frees = 2; b.AllocBytes = 64; b.FreeBytes = 0; b.Addresses = {0x3000};
  REQUIRE(r.SampleEvents({a, b}, meta) == 0);
  REQUIRE(sink.rows.size() == 3);  // a: inuse_objects + inuse_space; b: inuse_space only (allocs == frees)
  REQUIRE(sink.rows[0].kind == PA_KIND_MEM_INUSE_OBJECTS && sink.rows[0].value == 7 && sink.rows[1].kind == PA_KIND_MEM_INUSE_SPACE && sink.rows[1].value == 4000);
  REQUIRE(sink.rows[2].kind == PA_KIND_MEM_INUSE_SPACE && sink.rows[2].value == 64);
  for (auto& h : sink.rows) REQUIRE(h.hash_hi == 0 && h.hash_lo == 0 && h.pid == 4242 && h.tid == 4242 && h.timestamp_ns == 99);
  REQUIRE(sink.frames.size() == 3);  // 0x1000 interned once
  REQUIRE(sink.row_frames[0].size() == 3 && sink.row_frames[0][0] == sink.row_frames[0][2] && sink.row_frames[2].size() == 1);
  for (auto& f : sink.frames) {
    REQUIRE(f.kind == PA_FRAME_OOMPROF && sink.strings[f.type_name_sid] == "native");
    REQUIRE(sink.strings[f.function_name_sid] == "abcdef01" && sink.strings[f.source_file_sid] == "/srv/bin/victim");
  }
  REQUIRE(sink.label(sink.rows[0].labelset_id, "tenant") == "blue" && sink.label(sink.rows[0].labelset_id, "node") == "n");
  REQUIRE(r.memorySamples == 2 && r.ReportHostMetadataBlocking({}, 0, 0.0) == 0);
}

static void testFlushAndOfflineLog() {
  RecordingSink sink; Config cfg; cfg.nodeName = "n";
  uint64_t seen_rows = 0;
  cfg.onBatch = [&](const uint8_t*, uint64_t, uint64_t rows) { seen_rows += rows; };
  ParcaReporter r(&sink, cfg);
  REQUIRE(r.FlushOnce() == 0 && seen_rows == 0);  // empty interval is skipped (:1842-1845)
  Trace t; t.Hash = {1, 1}; TraceEventMeta m; m.PID = 1;
  r.ReportTraceEvent(&t, &m);
  REQUIRE(r.FlushOnce() == 1 && seen_rows == 1 && r.sampleWrites == 1);
  OfflineLog log;
  const uint8_t a[3] = {1, 2, 3}, b[2] = {9, 8};
  log.Append(a, 3); log.Append(b, 2);
  const uint8_t want[] = {0xA6, 0xE7, 0xCC, 0xCA, 0, 0, 0, 2, 0, 0, 0, 3, 1, 2, 3, 0, 0, 0, 2, 9, 8};
  REQUIRE(log.Bytes().size() == sizeof want && !memcmp(log.Bytes().data(), want, sizeof want) && log.Batches() == 2);
}

static int runGpu(const char* out_path) {  // TestSampleWriterV2_MultipleFrameTypes through the real C ABI
  pa_agg_config c;
  memset(&c, 0, sizeof c);
  c.abi_version = PA_ABI_VERSION; c.device = 0; c.hash_mode = PA_HASH_PROVIDED; c.samples_per_second = 19; c.max_samples = 1024; c.max_frames = 4096;
  c.label_flags = PA_LABEL_DISABLE_CPU | PA_LABEL_DISABLE_THREAD_ID | PA_LABEL_DISABLE_THREAD_COMM;
  pa_agg* agg = nullptr;
  if (pa_agg_create(&c, &agg) != PA_OK) { fprintf(stderr, "pa_agg_create failed\n"); return 2; }
  Sink* sink = NewAggSink(agg);
  Config cfg; cfg.nodeName = "";  // no node label: the labels struct stays empty like in the reference test
  std::vector<uint8_t> stream;
  cfg.onBatch = [&](const uint8_t* p, uint64_t n, uint64_t) { stream.assign(p, p + n); };
  {
    ParcaReporter r(sink, cfg);
    FileID fid{1, 2};
    r.ReportExecutable(ExecutableMetadata{fid, "/usr/bin/app", "build123"});
    Trace t1; t1.Hash = {1, 1}; t1.Frames = {nativeFrame(0x1000, true, fid)};
    Trace t2; t2.Hash = {2, 2}; t2.Frames = {kernelFrame(0x2000, "do_syscall_64", 100)};
    TraceEventMeta m; m.Timestamp = 1234567890; r.ReportTraceEvent(&t1, &m);
    m.Timestamp = 1234567891; r.ReportTraceEvent(&t2, &m);
    if (r.FlushOnce() != 2) { fprintf(stderr, "flush failed: %s\n", pa_agg_last_error(agg)); return 3; }
  }
  std::ofstream(out_path, std::ios::binary).write((const char*)stream.data(), (std::streamsize)stream.size());
  delete sink;
  pa_agg_destroy(agg);
  return stream.empty() ? 4 : 0;
}

// v1 schema, offline mode: two identical intervals -> .padata log with [samples, stacktraces(2 new), samples, stacktraces(0 new)]
static int runGpuV1(const char* out_path) {
  pa_agg_config c;
  memset(&c, 0, sizeof c);
  c.abi_version = PA_ABI_VERSION; c.device = 0; c.hash_mode = PA_HASH_PROVIDED; c.samples_per_second = 19; c.max_samples = 1024; c.max_frames = 4096;
  c.label_flags = PA_LABEL_DISABLE_CPU | PA_LABEL_DISABLE_THREAD_ID | PA_LABEL_DISABLE_THREAD_COMM;
  c.schema = PA_SCHEMA_V1;
  pa_agg* agg = nullptr;
  if (pa_agg_create(&c, &agg) != PA_OK) { fprintf(stderr, "pa_agg_create failed\n"); return 2; }
  Sink* sink = NewAggSink(agg);
  Config cfg; cfg.nodeName = ""; cfg.offlineV1Stacktraces = true;
  OfflineLog log;
  cfg.onBatch = [&](const uint8_t* p, uint64_t n, uint64_t) { log.Append(p, n); };
  {
    ParcaReporter r(sink, cfg);
    FileID fid{1, 2};
    r.ReportExecutable(ExecutableMetadata{fid, "/usr/bin/app", "build123"});
    Trace t1; t1.Hash = {1, 1}; t1.Frames = {nativeFrame(0x1000, true, fid)};
    Trace t2; t2.Hash = {2, 2}; t2.Frames = {kernelFrame(0x2000, "do_syscall_64", 100)};
    for (int interval = 0; interval < 2; interval++) {
      TraceEventMeta m; m.Timestamp = 1234567890; r.ReportTraceEvent(&t1, &m);
      m.Timestamp = 1234567891; r.ReportTraceEvent(&t2, &m);
      if (r.FlushOnce() != 2) { fprintf(stderr, "flush failed: %s\n", pa_agg_last_error(agg)); return 3; }
    }
    if (r.stacktraceWriteRequestBytes == 0) return 5;
  }
  std::ofstream(out_path, std::ios::binary).write((const char*)log.Bytes().data(), (std::streamsize)log.Bytes().size());
  delete sink;
  pa_agg_destroy(agg);
  return log.Batches() == 4 ? 0 : 4;
}

int main(int argc, char** argv) {
  if (argc >= 3 && !strcmp(argv[1], "--gpu")) return runGpu(argv[2]);
  if (argc >= 3 && !strcmp(argv[1], "--gpu-v1")) return runGpuV1(argv[2]);
  testMaybeFixTruncation();
  testLabelsAndInterning();
  testOrigins();
  testExecutables();
  testFlushAndOfflineLog();
  testOomprofSampleEvents();
  if (failures) { fprintf(stderr, "%d failure(s)\n", failures); return 1; }
  printf("ok\n");
  return 0;
}
