"""Tiny hand-built workloads that restate the reference's own unit tests as inputs.

reporter/arrow_v2_test.go and reporter/parca_reporter_test.go (parca-dev/parca-agent). The same
builders are used for the oracle KATs (CPU) and the CUDA-path KATs (GPU).
"""
import numpy as np

from parca_agent_b200 import abi, synth


def build(samples, frames, labelsets=None, label_flags=7, hash_mode=abi.PA_HASH_PROVIDED, external=(), strings=None):
    """samples: list of dict(hash=(hi,lo), frames=[frame idx], kind, ts, value, labelset, tid, cpu, comm)."""
    st = strings or synth.StringTable()
    fr = np.zeros(len(frames), dtype=abi.FRAME_DTYPE)
    for i, d in enumerate(frames):
        for k, v in d.items():
            fr[k][i] = st.sid(v) if k.endswith("_sid") and isinstance(v, (str, bytes)) else v
    ls = [sorted([(st.sid(k), st.sid(v)) for k, v in d.items()], key=lambda kv: st.strings[kv[0]]) for d in (labelsets or [{}])]
    hd = np.zeros(len(samples), dtype=abi.HDR_DTYPE)
    stream = []
    off = 0
    for i, s in enumerate(samples):
        hd["hash_hi"][i], hd["hash_lo"][i] = s.get("hash", (0, 0))
        hd["timestamp_ns"][i] = s.get("ts", 1234567890 + i)
        hd["value"][i] = s.get("value", 0)
        hd["kind"][i] = s.get("kind", abi.PA_KIND_CPU)
        hd["labelset_id"][i] = s.get("labelset", 0)
        hd["tid"][i] = s.get("tid", 0)
        hd["pid"][i] = s.get("pid", 0)
        hd["cpu"][i] = s.get("cpu", 0)
        hd["comm_sid"][i] = st.sid(s.get("comm", ""))
        hd["nframes"][i] = len(s["frames"])
        hd["frame_off"][i] = off
        stream.extend(s["frames"])
        off += len(s["frames"])
    ext = [(st.sid(k), st.sid(v)) for k, v in external]
    return synth.Workload("kat", st.strings, fr, ls, hd, _frame_ids=np.asarray(stream, dtype=np.uint64), hash_mode=hash_mode,
                          label_flags=label_flags, samples_per_second=19, external_labels=ext)


def native(addr, file="/usr/bin/test", build_id="abc123"):
    return dict(kind=abi.PA_FRAME_NATIVE, flags=3, type_name_sid="native", address_or_lineno=addr, exec_file_name_sid=file,
                exec_build_id_sid=build_id)


def kernel(addr, fn, line):
    return dict(kind=abi.PA_FRAME_KERNEL, flags=0, type_name_sid="kernel", address_or_lineno=addr, function_name_sid=fn, source_line=line)


def stack_dedup():
    """TestStacktraceDictBuilderV2_Deduplication (arrow_v2_test.go:144-204)."""
    frames = [native(0x1000), native(0x2000), native(0x3000)]
    s1 = dict(hash=(1, 2), frames=[0, 1])
    s2 = dict(hash=(3, 4), frames=[2])
    return build([s1, s1, s2, s1], frames)


def writer_basic():
    """TestSampleWriterV2_Basic (arrow_v2_test.go:206-255)."""
    return build([dict(hash=(1, 2), frames=[0], ts=1234567890)], [native(0x1000)], labelsets=[{"service": "my-service", "pod": "pod-1"}])


def multiple_frame_types():
    """TestSampleWriterV2_MultipleFrameTypes (arrow_v2_test.go:257-316)."""
    frames = [native(0x1000, "/usr/bin/app", "build123"), kernel(0x2000, "do_syscall_64", 100)]
    return build([dict(hash=(1, 1), frames=[0], ts=1234567890), dict(hash=(2, 2), frames=[1], ts=1234567891)], frames)


def func_dedup_in_stack():
    """TestFunctionDictBuilderV2_UsedInStacktrace (arrow_v2_test.go:318-364)."""
    frames = [kernel(0x1000, "do_syscall_64", 100), kernel(0x2000, "do_syscall_64", 200), kernel(0x3000, "sys_read", 50)]
    return build([dict(hash=(1, 1), frames=[0, 1, 2])], frames)


def null_lines():
    """TestStacktraceDictBuilderV2_NullLinesForUnsymbolizedFrames (arrow_v2_test.go:366-411)."""
    frames = [native(0x1000, "/usr/bin/app", "build123"), kernel(0x2000, "do_syscall_64", 100)]
    return build([dict(hash=(10, 20), frames=[0, 1])], frames)


def labels_cpu_sequence(flags=0):
    """TestLabelsForTID_* (parca_reporter_test.go:64-150): node label + per-sample cpu/thread_id/thread_name."""
    cpus = [0, 1, 0, 3, 2, 1, 3, 0]
    samples = [dict(hash=(1, 1), frames=[0], tid=4243, pid=4140, cpu=c, comm="myprocess") for c in cpus]
    return build(samples, [native(0x1000)], labelsets=[{"node": "test-node"}], label_flags=flags), cpus


# ---- inputs of the Go-side golden recipe (tools/golden/golden_parity_test.go feeds the SAME samples through the reference's
# ReportTraceEvent -> buildSampleRecordV2 -> ipc.Writer). Only mapping-less frames are used, so the Go side needs nothing from
# the un-vendored profiler module beyond libpf.Frame literals.
GO_PIN_TYPE_NAMES = {"native": "native", "kernel": "kernel", "python": "python"}  # libpf.FrameType.String(); overridden by tests/golden/go/frame_types.json


def interp(type_name, lineno, fn, file, line):
    return dict(kind=abi.PA_FRAME_INTERP, flags=0, type_name_sid=type_name, address_or_lineno=lineno, function_name_sid=fn, source_file_sid=file, source_line=line)


def go_pin_cases(type_names=None):
    """name -> workload; keep in lock-step with the cases in tools/golden/golden_parity_test.go"""
    tn = dict(GO_PIN_TYPE_NAMES, **(type_names or {}))
    nat = lambda addr: dict(kind=abi.PA_FRAME_NATIVE, flags=0, type_name_sid=tn["native"], address_or_lineno=addr)  # noqa: E731  (no mapping: "UNKNOWN", null build id)
    ker = lambda addr, fn, line: dict(kind=abi.PA_FRAME_KERNEL, flags=0, type_name_sid=tn["kernel"], address_or_lineno=addr, function_name_sid=fn, source_line=line)  # noqa: E731
    cases = {}
    # 1. one CPU sample, cached labels {pod, service}, per-sample labels disabled
    cases["basic"] = build([dict(hash=(1, 2), frames=[0], ts=1234567890, pid=100, tid=100)], [nat(0x1000)],
                           labelsets=[{"service": "my-service", "pod": "pod-1"}], label_flags=7)
    # 2. mixed frame kinds and origins, per-sample cpu / thread_id / thread_name labels, two pids with different label names
    #    (null back-fill), repeated stacks (ListView reuse), a function name longer than 12 bytes (StringView data block)
    frames = [nat(0x1000), nat(0x2000), ker(0xffffffff81000010, "do_syscall_64", 100), ker(0xffffffff81000020, "", 0),
              interp(tn["python"], 10, "handler", "/srv/app/main.py", 42), interp(tn["python"], 11, "a_function_name_longer_than_twelve_bytes", "", 7),
              interp(tn["python"], 12, "", "ignored.py", 9)]
    ls = [{"node": "test-node"}, {"node": "test-node", "job": "batch"}]
    S, O, C = abi.PA_KIND_CPU, abi.PA_KIND_OFFCPU, abi.PA_KIND_CUDA
    rows = [
        dict(hash=(1, 1), frames=[0, 2], kind=S, ts=1000, pid=10, tid=11, cpu=0, comm="alpha", labelset=0),
        dict(hash=(1, 1), frames=[0, 2], kind=S, ts=1001, pid=10, tid=11, cpu=0, comm="alpha", labelset=0),
        dict(hash=(2, 2), frames=[4, 5, 1], kind=S, ts=1002, pid=20, tid=21, cpu=3, comm="beta", labelset=1),
        dict(hash=(3, 3), frames=[3, 6], kind=O, ts=1003, value=5000, pid=20, tid=22, cpu=3, comm="", labelset=1),
        dict(hash=(1, 1), frames=[0, 2], kind=C, ts=1004, value=777, pid=10, tid=11, cpu=1, comm="alpha", labelset=0),
        dict(hash=(4, 4), frames=[], kind=S, ts=1005, pid=10, tid=12, cpu=1, comm="alpha", labelset=0),
    ]
    cases["mixed"] = build(rows, frames, labelsets=ls, label_flags=0)
    # 3. the same rows with external labels: a new name and one that collides with a sample label
    cases["mixed_external"] = build(rows, frames, labelsets=ls, label_flags=0, external=[("cluster", "prod"), ("job", "ext-job")])
    return cases
