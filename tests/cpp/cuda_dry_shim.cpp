// cuda_dry_shim.cpp — TEST INFRASTRUCTURE ONLY (tests/test_host_dry_run.py).
//
// A stand-in for the handful of CUDA runtime entry points libdcvc_b200 uses, LD_PRELOADed in front of a copy of the
// library that the test relinks against the *shared* runtime (the shipped library links the static one and cannot be
// interposed).  "Device" memory is zeroed host memory, copies are memcpy, streams are synchronous, a captured graph
// is the list of its launches.  A kernel launch is handed to the host-side restatement of that kernel in
// cuda_emu_kernels.cpp (DRY_SHIM_EMULATE=1), or only counted (default: nothing is computed).  The point is to drive
// the host side of the codecs — parameter loading and weight repacking, arena sizing, the plan of every GEMM (geometry
// checks, tile plans, tensor-map encoding with the driver's documented argument rules), operand wiring, segment
// building, the compress / decompress control flow and the rANS hand-off — on a machine without a GPU, so that shape
// and wiring mistakes surface in the CPU test tier.  A launch of a kernel the emulation does not know fails loudly.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <vector>

// cuda_emu_kernels.cpp
extern "C" int emu_lookup(const char* name);
extern "C" int emu_num_args(int k);
extern "C" int emu_arg_size(int k, int i);
extern "C" int emu_run(int k, void** args);
extern "C" void (*emu_access_hook)(const void* lo, const void* hi, int is_write);
extern "C" int emu_reported;
extern "C" unsigned long long emu_map_magic();

extern "C" {

typedef int cudaError_t;
typedef struct CUstream_st* cudaStream_t;
typedef struct CUevent_st* cudaEvent_t;
typedef struct CUgraph_st* cudaGraph_t;
typedef struct CUgraphExec_st* cudaGraphExec_t;
struct dim3 { unsigned x, y, z; };

static std::atomic<long long> g_launches{0}, g_graph_launches{0}, g_maps{0}, g_bytes{0}, g_forks{0};
static int g_fail_map = 0;

long long dry_shim_launches() { return g_launches.load(); }
long long dry_shim_graph_launches() { return g_graph_launches.load(); }
long long dry_shim_capture_forks() { return g_forks.load(); }   // streams that joined a capture by waiting for its event
long long dry_shim_tensor_maps() { return g_maps.load(); }
long long dry_shim_bytes() { return g_bytes.load(); }

// ---- fat binary registration (no driver is ever touched) --------------------------------------------------------
void** __cudaRegisterFatBinary(void*) { static void* h[4]; return h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
static std::map<const void*, std::string>& kernel_names() { static std::map<const void*, std::string> m; return m; }
void __cudaRegisterFunction(void**, const char* host_fun, char*, const char* device_name, int, void*, void*, void*, void*, int*)
{
    kernel_names()[host_fun] = device_name;
}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}

static thread_local struct { dim3 g, b; size_t smem; void* st; } t_cfg;
unsigned __cudaPushCallConfiguration(dim3 g, dim3 b, size_t smem, void* st) { t_cfg = { g, b, smem, st }; return 0; }
cudaError_t __cudaPopCallConfiguration(dim3* g, dim3* b, size_t* smem, void* st)
{
    *g = t_cfg.g; *b = t_cfg.b; *smem = t_cfg.smem; *static_cast<void**>(st) = t_cfg.st;
    return 0;
}

// ---- device / memory ---------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
cudaError_t cudaSetDevice(int) { return 0; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
cudaError_t cudaDeviceGetAttribute(int* v, int attr, int)
{
    *v = (attr == 16 /* cudaDevAttrMultiProcessorCount */) ? 148 : 0;
    return 0;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return 0; }
cudaError_t cudaDeviceSynchronize() { return 0; }
static thread_local cudaError_t t_last_error = 0;
cudaError_t cudaGetLastError() { const cudaError_t e = t_last_error; t_last_error = 0; return e; }
const char* cudaGetErrorString(cudaError_t) { return "dry-run shim"; }

static cudaError_t alloc_zeroed(void** p, size_t n)
{
    void* q = nullptr;
    if (posix_memalign(&q, 1024, n ? n : 1)) return 2;
    memset(q, 0, n);
    g_bytes += static_cast<long long>(n);
    *p = q;
    return 0;
}
cudaError_t cudaMalloc(void** p, size_t n) { return alloc_zeroed(p, n); }
cudaError_t cudaFree(void* p) { free(p); return 0; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return alloc_zeroed(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memmove(d, s, n); return 0; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return 0; }
cudaError_t dry_memset_async(void* d, int v, size_t n, cudaStream_t st);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) { return dry_memset_async(d, v, n, st); }

// ---- streams / events / graphs -------------------------------------------------------------------------------------
static void* token() { return malloc(8); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = static_cast<cudaStream_t>(token()); return 0; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = static_cast<cudaStream_t>(token()); return 0; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return 0; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = static_cast<cudaEvent_t>(token()); return 0; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = static_cast<cudaEvent_t>(token()); return 0; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return 0; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.25f; return 0; }   // a made-up, non-zero interval
// a captured graph = the launches (kernel + a private copy of its arguments) and memsets issued while capturing
struct GraphOp {
    int kernel = -1;                       // emulation table index, -1: memset
    std::vector<std::shared_ptr<void>> args;   // 64-byte aligned private copies (PwGemmParams is alignas(64))
    void* dst = nullptr; int value = 0; size_t bytes = 0;
    cudaStream_t lane = nullptr;           // the stream the op was captured on (parallel branch of the graph)
    int lane_i = 0;                        // its index in the capture (0 = origin stream)
    std::vector<int> vc;                   // vector clock of the op (happens-before between branches)
};
struct Graph { std::vector<GraphOp> ops; bool multi_lane = false, lanes_checked = false; };
static std::map<cudaStream_t, Graph*>& capturing() { static std::map<cudaStream_t, Graph*> m; return m; }
// Cross-stream capture (fork / join), as the runtime does it: an event recorded on a capturing stream carries the
// capture; a stream that waits for it joins the capture (its launches land in the same graph); every forked stream must
// have been joined back into the origin stream (record on the fork, wait on the origin) when the capture ends
// (cudaErrorStreamCaptureUnjoined otherwise).
struct CaptureState {
    cudaStream_t origin;
    std::map<cudaStream_t, int> lane_idx;      // origin = 0, forks in the order they joined the capture
    std::vector<std::vector<int>> vc;          // vector clock per lane: what of every lane's work it has seen
};
struct EventSnap { Graph* g; cudaStream_t st; std::vector<int> vc; };
static std::map<Graph*, CaptureState>& capture_state() { static std::map<Graph*, CaptureState> m; return m; }
static std::map<cudaEvent_t, EventSnap>& event_capture() { static std::map<cudaEvent_t, EventSnap> m; return m; }
static int vc_get(const std::vector<int>& v, int i) { return i < static_cast<int>(v.size()) ? v[i] : 0; }
static void vc_merge(std::vector<int>& into, const std::vector<int>& from)
{
    if (into.size() < from.size()) into.resize(from.size(), 0);
    for (size_t i = 0; i < from.size(); ++i) if (from[i] > into[i]) into[i] = from[i];
}
// stamps an op captured on `st` (a kernel or a memset) with its lane and clock
static void stamp_op(Graph* g, cudaStream_t st, GraphOp& op)
{
    CaptureState& cs = capture_state()[g];
    const int l = cs.lane_idx[st];
    std::vector<int>& v = cs.vc[l];
    if (static_cast<int>(v.size()) <= l) v.resize(l + 1, 0);
    ++v[l];
    op.lane = st;
    op.lane_i = l;
    op.vc = v;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st)
{
    auto cap = capturing().find(st);
    if (cap == capturing().end()) { event_capture().erase(e); return 0; }
    CaptureState& cs = capture_state()[cap->second];
    event_capture()[e] = EventSnap{ cap->second, st, cs.vc[cs.lane_idx[st]] };
    return 0;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned)
{
    auto ev = event_capture().find(e);
    if (ev == event_capture().end()) return 0;
    Graph* g = ev->second.g;
    auto csi = capture_state().find(g);
    if (csi == capture_state().end()) return 0;               // capture already over: a plain (completed) event
    CaptureState& cs = csi->second;
    auto cap = capturing().find(st);
    if (cap == capturing().end()) {                           // fork: the stream joins the capture behind the event
        capturing()[st] = g;
        const int l = static_cast<int>(cs.vc.size());
        cs.lane_idx[st] = l;
        cs.vc.push_back(ev->second.vc);
        ++g_forks;
        return 0;
    }
    if (cap->second != g) return 905;                         // cudaErrorStreamCaptureMerge
    vc_merge(cs.vc[cs.lane_idx[st]], ev->second.vc);          // st now runs behind everything the event had seen
    return 0;
}
static const bool g_emulate = []() { const char* e = getenv("DRY_SHIM_EMULATE"); return e && e[0] == '1'; }();

static cudaError_t run_op(const GraphOp& op)
{
    if (op.kernel < 0) { memset(op.dst, op.value, op.bytes); return 0; }
    if (!g_emulate) return 0;
    std::vector<void*> ptrs(op.args.size());
    for (size_t i = 0; i < op.args.size(); ++i) ptrs[i] = op.args[i].get();
    return emu_run(op.kernel, ptrs.data()) ? 719 /* launch failure */ : 0;
}
static cudaError_t submit_kernel(const void* fn, void** args, cudaStream_t st)
{
    ++g_launches;
    auto it = kernel_names().find(fn);
    if (it == kernel_names().end()) { fprintf(stderr, "dry shim: launch of an unregistered kernel\n"); return 98; }
    const int k = emu_lookup(it->second.c_str());
    if (k < 0) {
        if (!g_emulate) return 0;
        fprintf(stderr, "dry shim: no host restatement of kernel %s\n", it->second.c_str());
        return 98;  // invalid device function
    }
    GraphOp op;
    op.kernel = k;
    for (int i = 0; i < emu_num_args(k); ++i) {
        const size_t n = static_cast<size_t>(emu_arg_size(k, i));
        void* q = nullptr;
        if (posix_memalign(&q, 64, (n + 63) & ~static_cast<size_t>(63))) return 2;
        memcpy(q, args[i], n);
        op.args.emplace_back(q, free);
    }
    auto cap = capturing().find(st);
    if (cap != capturing().end()) { stamp_op(cap->second, st, op); cap->second->ops.push_back(std::move(op)); return 0; }
    return run_op(op);
}

// Lane race check.  The library's multi-lane segments fork every lane at the segment's start and join at its end, so
// ops captured on different streams of one graph may run concurrently on the device, while this shim runs them one
// after the other.  On the first (emulated) launch of such a graph every kernel reports the byte ranges it reads and
// writes; a range written on one lane that overlaps any range touched on another lane fails the launch.
struct Access { const uint8_t *lo, *hi; int write; size_t op; };
static std::vector<Access>* g_collect = nullptr;
static size_t g_collect_op = 0;
static void collect_access(const void* lo, const void* hi, int is_write)
{
    if (g_collect) g_collect->push_back({ static_cast<const uint8_t*>(lo), static_cast<const uint8_t*>(hi), is_write, g_collect_op });
}
static bool happens_before(const GraphOp& a, const GraphOp& b) { return vc_get(a.vc, a.lane_i) <= vc_get(b.vc, a.lane_i); }
static cudaError_t run_graph_checked(Graph* g)
{
    std::vector<Access> acc;
    std::vector<size_t> silent;  // kernels whose restatement reports no accesses
    g_collect = &acc;
    emu_access_hook = collect_access;
    cudaError_t rc = 0;
    for (size_t i = 0; i < g->ops.size(); ++i) {
        const GraphOp& op = g->ops[i];
        g_collect_op = i;
        emu_reported = 0;
        rc = run_op(op);
        if (rc) break;
        if (op.kernel >= 0 && !emu_reported) silent.push_back(i);
        if (op.kernel < 0) acc.push_back({ static_cast<const uint8_t*>(op.dst), static_cast<const uint8_t*>(op.dst) + op.bytes, 1, i });
    }
    emu_access_hook = nullptr;
    g_collect = nullptr;
    if (rc) return rc;
    // a kernel that reports nothing can only be vouched for when it cannot overlap another branch at all, i.e. when it is
    // ordered against every op of every other lane (ops in front of the fork or behind the join)
    for (size_t i : silent) {
        const GraphOp& a = g->ops[i];
        for (const GraphOp& b : g->ops) {
            if (b.lane_i == a.lane_i || happens_before(a, b) || happens_before(b, a)) continue;
            fprintf(stderr, "dry shim: op %zu of a multi-lane graph may run beside another branch but does not report its accesses "
                            "(no race check possible)\n", i);
            return 719;
        }
    }
    // two ops race when their ranges overlap, at least one writes, they sit on different branches and neither is
    // ordered behind the other by the capture's event edges
    for (size_t i = 0; i < acc.size(); ++i) {
        if (!acc[i].write) continue;
        const GraphOp& a = g->ops[acc[i].op];
        for (size_t j = 0; j < acc.size(); ++j) {
            if (acc[j].op == acc[i].op) continue;
            const GraphOp& b = g->ops[acc[j].op];
            if (a.lane_i == b.lane_i) continue;
            if (!(acc[i].lo < acc[j].hi && acc[j].lo < acc[i].hi)) continue;
            if (happens_before(a, b) || happens_before(b, a)) continue;
            fprintf(stderr, "dry shim: lane race: op %zu (lane %d) writes [%p, %p), op %zu (lane %d) %s [%p, %p), and no event edge orders them\n",
                    acc[i].op, a.lane_i, (const void*)acc[i].lo, (const void*)acc[i].hi, acc[j].op, b.lane_i,
                    acc[j].write ? "writes" : "reads", (const void*)acc[j].lo, (const void*)acc[j].hi);
            return 719;
        }
    }
    return 0;
}

cudaError_t cudaStreamBeginCapture(cudaStream_t s, int)
{
    if (capturing().count(s)) return 900;
    Graph* g = new Graph();
    capturing()[s] = g;
    CaptureState cs;
    cs.origin = s;
    cs.lane_idx[s] = 0;
    cs.vc.push_back(std::vector<int>(1, 0));
    capture_state()[g] = cs;
    return 0;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g)
{
    auto it = capturing().find(s);
    if (it == capturing().end()) return 901;
    Graph* gr = it->second;
    auto cs = capture_state().find(gr);
    if (cs == capture_state().end() || cs->second.origin != s) return 902;   // cudaErrorStreamCaptureUnmatched
    bool unjoined = false;
    gr->multi_lane = cs->second.lane_idx.size() > 1;
    for (auto& f : cs->second.lane_idx) {
        if (f.second == 0) continue;
        // joined = the origin stream has seen all of the fork's work (record on the fork, wait on the origin)
        const int own = vc_get(cs->second.vc[f.second], f.second);
        unjoined = unjoined || vc_get(cs->second.vc[0], f.second) < own;
        capturing().erase(f.first);
    }
    capture_state().erase(cs);
    capturing().erase(it);
    if (unjoined) { delete gr; *g = nullptr; fprintf(stderr, "dry shim: capture ended with an unjoined forked stream\n"); return 904; }
    *g = reinterpret_cast<cudaGraph_t>(gr);
    return 0;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long)
{
    *e = reinterpret_cast<cudaGraphExec_t>(new Graph(*reinterpret_cast<Graph*>(g)));
    return 0;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t)
{
    ++g_graph_launches;
    Graph* gr = reinterpret_cast<Graph*>(e);
    if (g_emulate && gr->multi_lane && !gr->lanes_checked) {
        gr->lanes_checked = true;
        return run_graph_checked(gr);
    }
    for (const GraphOp& op : gr->ops) {
        const cudaError_t r = run_op(op);
        if (r) return r;
    }
    return 0;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete reinterpret_cast<Graph*>(g); return 0; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete reinterpret_cast<Graph*>(e); return 0; }

// ---- launches ------------------------------------------------------------------------------------------------------
struct LaunchConfigHead { dim3 grid, block; size_t smem; cudaStream_t stream; void* attrs; unsigned num_attrs; };
static cudaError_t check_dims(const dim3& g, const dim3& b, size_t smem)
{
    if (g.x == 0 || g.y == 0 || g.z == 0 || b.x == 0 || b.y == 0 || b.z == 0) return 9;       // invalid configuration
    if (static_cast<unsigned long long>(b.x) * b.y * b.z > 1024 || smem > 232448) return 9;
    if (g.y > 65535 || g.z > 65535) return 9;
    return 0;
}
cudaError_t dry_memset_async(void* d, int v, size_t n, cudaStream_t st)
{
    auto cap = capturing().find(st);
    if (cap == capturing().end()) { memset(d, v, n); return 0; }
    GraphOp op;
    op.dst = d; op.value = v; op.bytes = n;
    stamp_op(cap->second, st, op);
    cap->second->ops.push_back(op);
    return 0;
}
cudaError_t cudaLaunchKernel(const void* fn, dim3 g, dim3 b, void** args, size_t smem, cudaStream_t st)
{
    cudaError_t e = check_dims(g, b, smem);
    if (!e) e = submit_kernel(fn, args, st);
    if (e) t_last_error = e;      // a failed launch is also what cudaGetLastError() reports next (<<< >>> launches)
    return e;
}
cudaError_t cudaLaunchKernelExC(const void* cfg, const void* fn, void** args)
{
    const LaunchConfigHead* c = static_cast<const LaunchConfigHead*>(cfg);
    cudaError_t e = check_dims(c->grid, c->block, c->smem);
    if (!e) e = submit_kernel(fn, args, c->stream);
    if (e) t_last_error = e;
    return e;
}
cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
cudaError_t cudaOccupancyMaxActiveClusters(int* n, const void*, const void*)
{
    *n = 0;
    return 1;   // "not available": the library falls back to num_sms / cluster size
}

// ---- cuTensorMapEncodeTiled with the argument rules of the driver API documentation ------------------------------------
static int fake_encode_tiled(void* map, int dtype, unsigned rank, void* addr, const uint64_t* dims, const uint64_t* strides,
                             const uint32_t* box, const uint32_t* estr, int interleave, int swizzle, int, int)
{
    ++g_maps;
    if (g_fail_map) return 1;
    if (!map || rank < 1 || rank > 5) return 1;
    if (reinterpret_cast<uintptr_t>(addr) & 15) return 1;
    const int esize = (dtype == 6 /* FLOAT16 */) ? 2 : 0;
    if (!esize || interleave != 0) return 1;
    for (unsigned i = 0; i < rank; ++i) {
        if (dims[i] == 0 || dims[i] > (1ull << 32)) return 1;
        if (box[i] == 0 || box[i] > 256) return 1;
        if (estr[i] == 0 || estr[i] > 8) return 1;
    }
    for (unsigned i = 0; i + 1 < rank; ++i)
        if ((strides[i] & 15) || strides[i] >= (1ull << 40)) return 1;
    const uint64_t inner = static_cast<uint64_t>(box[0]) * esize;
    if (inner & 15) return 1;
    const uint64_t span = swizzle == 1 ? 32 : swizzle == 2 ? 64 : swizzle == 3 ? 128 : (1ull << 40);
    if (inner > span) return 1;
    // transparent map for the host restatement of the kernels (FakeMap in cuda_emu_kernels.cpp)
    struct { uint64_t magic; const void* ptr; uint32_t rank, swz; uint64_t dims[5]; uint64_t strides[4]; uint32_t box[5]; } f;
    memset(&f, 0, sizeof(f));
    f.magic = emu_map_magic(); f.ptr = addr; f.rank = rank; f.swz = static_cast<uint32_t>(swizzle);
    for (unsigned i = 0; i < rank; ++i) { f.dims[i] = dims[i]; f.box[i] = box[i]; }
    for (unsigned i = 0; i + 1 < rank; ++i) f.strides[i] = strides[i];
    memset(map, 0, 128);
    memcpy(map, &f, sizeof(f));
    return 0;
}
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, int* qres)
{
    if (strcmp(name, "cuTensorMapEncodeTiled") != 0) { *fn = nullptr; if (qres) *qres = 1; return 0; }
    *fn = reinterpret_cast<void*>(&fake_encode_tiled);
    if (qres) *qres = 0;
    return 0;
}

}  // extern "C"
