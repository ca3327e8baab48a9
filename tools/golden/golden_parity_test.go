// golden_parity_test.go — produces Go-side golden bytes for the B200 port's parity tests.
//
// NOT compiled in this repository (there is no Go toolchain in the build image). Copy it into the reference checkout's
// reporter/ directory (package reporter: it needs the unexported ReportTraceEvent plumbing) and run
//
//	PA_GOLDEN_DIR=/path/to/repo/tests/golden/go go test ./reporter -run TestWriteB200Golden -count=1
//
// It feeds the inputs of tests/kat_workloads.py::go_pin_cases through the reference's own
// ReportTraceEvent (parca_reporter.go:219) -> buildSampleRecordV2 (:1742) -> ipc.NewWriter(WithSchema, WithAllocator)
// (:1779-1790, the offline-mode V2 bytes) and writes, per case, <name>.arrows (the IPC stream) and <name>.padata (the
// offline log framing of setupOfflineModeLog / logDataForOfflineModeV2), plus frame_types.json with the
// libpf.FrameType.String() values the cases used. tests/test_go_golden.py then requires the CPU oracle (and, on a GPU
// box, the CUDA path) to reproduce those files byte for byte.
//
// Only mapping-less frames are used, so nothing but libpf.Frame literals is needed from the profiler module.
package reporter

import (
	"bytes"
	"context"
	"encoding/binary"
	"encoding/json"
	"os"
	"path/filepath"
	"testing"
	"time"

	"github.com/apache/arrow-go/v18/arrow/ipc"
	"github.com/apache/arrow-go/v18/arrow/memory"
	lru "github.com/elastic/go-freelru"
	"github.com/prometheus/client_golang/prometheus"
	"github.com/prometheus/prometheus/model/labels"
	"github.com/stretchr/testify/require"
	"go.opentelemetry.io/ebpf-profiler/libpf"
	"go.opentelemetry.io/ebpf-profiler/reporter/samples"
	"go.opentelemetry.io/ebpf-profiler/support"

	"github.com/parca-dev/parca-agent/metadata"
)

type goldenSample struct {
	hashHi, hashLo uint64
	frames         []int
	origin         libpf.Origin // support.TraceOriginSampling / OffCPU / Cuda
	ts             int64
	value          int64
	pid, tid       libpf.PID
	cpu            int
	comm           string
}

func goldenReporter(t *testing.T, disableCPU, disableTID, disableComm bool, external []Label) *ParcaReporter {
	t.Helper()
	mem := memory.NewGoAllocator()
	lbls, err := lru.NewSynced[libpf.PID, labelRetrievalResult](1024, libpf.PID.Hash32)
	require.NoError(t, err)
	lbls.SetLifetime(10 * time.Minute)
	stacks, err := lru.NewSynced[libpf.TraceHash, libpf.Frames](1024, libpf.TraceHash.Hash32)
	require.NoError(t, err)
	execs, err := lru.NewSynced[libpf.FileID, metadata.ExecInfo](1024, libpf.FileID.Hash32)
	require.NoError(t, err)
	return &ParcaReporter{
		labels:                 lbls,
		stacks:                 stacks,
		executables:            execs,
		mem:                    mem,
		nodeName:               "test-node",
		samplesPerSecond:       19,
		useV2Schema:            true,
		sampleWriterV2:         NewSampleWriterV2(mem),
		externalLabels:         external,
		disableCPULabel:        disableCPU,
		disableThreadIDLabel:   disableTID,
		disableThreadCommLabel: disableComm,
		emptySamples:           prometheus.NewCounter(prometheus.CounterOpts{Name: "golden_empty_samples"}),
		skippedByRelabeling:    prometheus.NewCounter(prometheus.CounterOpts{Name: "golden_skipped"}),
	}
}

// writeGolden: everything after the samples were reported — exactly the serialisation of logDataForOfflineModeV2.
func writeGolden(t *testing.T, r *ParcaReporter, dir, name string) {
	t.Helper()
	record := r.buildSampleRecordV2(context.Background())
	defer record.Release()
	var buf bytes.Buffer
	w := ipc.NewWriter(&buf, ipc.WithSchema(record.Schema()), ipc.WithAllocator(r.mem))
	require.NoError(t, w.Write(record))
	require.NoError(t, w.Close())
	require.NoError(t, os.WriteFile(filepath.Join(dir, name+".arrows"), buf.Bytes(), 0o644))
	// offline log framing: magic, u16 version, u16 batch count, then u32 BE size + stream (parca_reporter.go:1102-1116, :1807-1831)
	var log bytes.Buffer
	log.Write([]byte{0xA6, 0xE7, 0xCC, 0xCA, 0, 0, 0, 1})
	require.NoError(t, binary.Write(&log, binary.BigEndian, uint32(buf.Len())))
	log.Write(buf.Bytes())
	require.NoError(t, os.WriteFile(filepath.Join(dir, name+".padata"), log.Bytes(), 0o644))
}

func report(t *testing.T, r *ParcaReporter, frames []libpf.Frame, s goldenSample) {
	t.Helper()
	fs := libpf.Frames{}
	for _, i := range s.frames {
		f := frames[i]
		fs.Append(&f)
	}
	trace := &libpf.Trace{Hash: libpf.NewTraceHash(s.hashHi, s.hashLo), Frames: fs}
	meta := &samples.TraceEventMeta{
		Timestamp: libpf.UnixTime64(s.ts),
		Comm:      libpf.Intern(s.comm),
		PID:       s.pid,
		TID:       s.tid,
		CPU:       s.cpu,
		Origin:    s.origin,
		OffTime:   s.value,
	}
	require.NoError(t, r.ReportTraceEvent(trace, meta))
}

func TestWriteB200Golden(t *testing.T) {
	dir := os.Getenv("PA_GOLDEN_DIR")
	if dir == "" {
		t.Skip("PA_GOLDEN_DIR not set")
	}
	require.NoError(t, os.MkdirAll(dir, 0o755))
	types, err := json.Marshal(map[string]string{
		"native": libpf.NativeFrame.String(), "kernel": libpf.KernelFrame.String(), "python": libpf.PythonFrame.String(),
	})
	require.NoError(t, err)
	require.NoError(t, os.WriteFile(filepath.Join(dir, "frame_types.json"), types, 0o644))

	// ---- case "basic" (kat_workloads.go_pin_cases()["basic"])
	{
		r := goldenReporter(t, true, true, true, nil)
		r.labels.Add(libpf.PID(100), labelRetrievalResult{labels: labels.FromStrings("pod", "pod-1", "service", "my-service"), keep: true})
		frames := []libpf.Frame{{Type: libpf.NativeFrame, AddressOrLineno: 0x1000}}
		report(t, r, frames, goldenSample{hashHi: 1, hashLo: 2, frames: []int{0}, origin: support.TraceOriginSampling, ts: 1234567890, pid: 100, tid: 100})
		writeGolden(t, r, dir, "basic")
	}

	// ---- cases "mixed" / "mixed_external"
	frames := []libpf.Frame{
		{Type: libpf.NativeFrame, AddressOrLineno: 0x1000},
		{Type: libpf.NativeFrame, AddressOrLineno: 0x2000},
		{Type: libpf.KernelFrame, AddressOrLineno: 0xffffffff81000010, FunctionName: libpf.Intern("do_syscall_64"), SourceLine: 100},
		{Type: libpf.KernelFrame, AddressOrLineno: 0xffffffff81000020},
		{Type: libpf.PythonFrame, AddressOrLineno: 10, FunctionName: libpf.Intern("handler"), SourceFile: libpf.Intern("/srv/app/main.py"), SourceLine: 42},
		{Type: libpf.PythonFrame, AddressOrLineno: 11, FunctionName: libpf.Intern("a_function_name_longer_than_twelve_bytes"), SourceLine: 7},
		{Type: libpf.PythonFrame, AddressOrLineno: 12, SourceFile: libpf.Intern("ignored.py"), SourceLine: 9},
	}
	rows := []goldenSample{
		{1, 1, []int{0, 2}, support.TraceOriginSampling, 1000, 0, 10, 11, 0, "alpha"},
		{1, 1, []int{0, 2}, support.TraceOriginSampling, 1001, 0, 10, 11, 0, "alpha"},
		{2, 2, []int{4, 5, 1}, support.TraceOriginSampling, 1002, 0, 20, 21, 3, "beta"},
		{3, 3, []int{3, 6}, support.TraceOriginOffCPU, 1003, 5000, 20, 22, 3, ""},
		{1, 1, []int{0, 2}, support.TraceOriginCuda, 1004, 777, 10, 11, 1, "alpha"},
		{4, 4, []int{}, support.TraceOriginSampling, 1005, 0, 10, 12, 1, "alpha"},
	}
	for _, c := range []struct {
		name string
		ext  []Label
	}{
		{"mixed", nil},
		{"mixed_external", []Label{{Name: "cluster", Value: "prod"}, {Name: "job", Value: "ext-job"}}},
	} {
		r := goldenReporter(t, false, false, false, c.ext)
		r.labels.Add(libpf.PID(10), labelRetrievalResult{labels: labels.FromStrings("node", "test-node"), keep: true})
		r.labels.Add(libpf.PID(20), labelRetrievalResult{labels: labels.FromStrings("job", "batch", "node", "test-node"), keep: true})
		for _, s := range rows {
			report(t, r, frames, s)
		}
		writeGolden(t, r, dir, c.name)
	}
}
