// b2a_api.cu -- the C ABI of include/b2aruco.h: handles, device memory, stream, launch order.
// Everything that computes is a kernel in detect_kernels.cuh / ekf_kernels.cuh or the pose
// kernel below; this file only orchestrates.  No CPU fallback: without a CUDA device the
// create calls fail.
#include "../../include/b2aruco.h"
#include "detect_kernels.cuh"
#include "board_core.h"
#include "ekf_kernels.cuh"
#include "pose_core.h"
#include "draw_core.h"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <queue>
#include <unordered_map>
#include <string>
#include <thread>
#include <vector>

using namespace b2a;

static thread_local std::string g_err;
static int set_err(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return set_err(B2A_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));      \
    } while (0)

extern "C" const char *b2a_last_error(void) { return g_err.c_str(); }
extern "C" const char *b2a_version(void) { return "b2aruco 0.1 (sm_100a)"; }

extern "C" int b2a_host_alloc(size_t bytes, int write_combined, void **out)
{
    if (!out || bytes == 0) return set_err(B2A_ERR_INVALID, "null argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0));
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(B2A_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
    return B2A_OK;
}
extern "C" void b2a_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" void b2a_default_detector_params(b2a_detector_params *p)
{
    p->adaptiveThreshWinSizeMin = 3; p->adaptiveThreshWinSizeMax = 23; p->adaptiveThreshWinSizeStep = 10;
    p->adaptiveThreshConstant = 7.0;
    p->minMarkerPerimeterRate = 0.03; p->maxMarkerPerimeterRate = 4.0;
    p->polygonalApproxAccuracyRate = 0.03; p->minCornerDistanceRate = 0.05;
    p->minDistanceToBorder = 3; p->minMarkerDistanceRate = 0.125; p->minGroupDistance = 0.21f;
    p->markerBorderBits = 1; p->perspectiveRemovePixelPerCell = 4; p->perspectiveRemoveIgnoredMarginPerCell = 0.13;
    p->maxErroneousBitsInBorderRate = 0.35; p->minOtsuStdDev = 5.0; p->errorCorrectionRate = 0.6;
    p->cornerRefinementMethod = 0; p->cornerRefinementWinSize = 5; p->relativeCornerRefinmentWinSize = 0.3;
    p->cornerRefinementMaxIterations = 30; p->cornerRefinementMinAccuracy = 0.1; p->detectInvertedMarker = 0;
    p->useAruco3Detection = 0; p->minSideLengthCanonicalImg = 32; p->minMarkerLengthRatioOriginalImg = 0.f;
}

// ------------------------------------------------------------------------------------------------
// predefined dictionaries (tables generated from data/dict_tables.inc)
// ------------------------------------------------------------------------------------------------
namespace {
struct Family { const char *name; int markerSize, nMarkers, nBytes; const char *hex; };
struct Predef { int id; const char *name; const char *family; int nMarkers, maxCorr; };
#define B2A_DICT_FAMILY(name, ms, n, nb, hex) {#name, ms, n, nb, hex},
#define B2A_DICT_PREDEFINED(id, name, fam, n, mc)
static const Family k_families[] = {
#include "../data/dict_tables.inc"
};
#undef B2A_DICT_FAMILY
#undef B2A_DICT_PREDEFINED
#define B2A_DICT_FAMILY(name, ms, n, nb, hex)
#define B2A_DICT_PREDEFINED(id, name, fam, n, mc) {id, #name, #fam, n, mc},
static const Predef k_predef[] = {
#include "../data/dict_tables.inc"
};
#undef B2A_DICT_FAMILY
#undef B2A_DICT_PREDEFINED

static int hexval(char c) { return c <= '9' ? c - '0' : c - 'a' + 10; }

// rotation r = bit matrix rotated counter-clockwise r times (np.rot90(bits, r))
static std::vector<uint8_t> build_table(const Family &f, int n)
{
    const int ms = f.markerSize, nb = f.nBytes, nbits = ms * ms;
    std::vector<uint8_t> t((size_t)n * 4 * nb, 0);
    std::vector<uint8_t> bits(nbits), rot(nbits);
    auto shift_of = [&](int i) { const int byte = i / 8; return (byte == nb - 1 && (nbits % 8)) ? (nbits % 8) - 1 - (i % 8) : 7 - (i % 8); };
    for (int m = 0; m < n; ++m) {
        for (int i = 0; i < nbits; ++i) {
            const int byte = i / 8;
            const int v = hexval(f.hex[((size_t)m * nb + byte) * 2]) * 16 + hexval(f.hex[((size_t)m * nb + byte) * 2 + 1]);
            bits[i] = (v >> shift_of(i)) & 1;
        }
        for (int r = 0; r < 4; ++r) {
            uint8_t *o = &t[((size_t)m * 4 + r) * nb];
            for (int i = 0; i < nbits; ++i) o[i / 8] |= (uint8_t)(bits[i] << shift_of(i));
            // rot90 ccw: out[y][x] = in[x][ms-1-y]
            for (int y = 0; y < ms; ++y) for (int x = 0; x < ms; ++x) rot[y * ms + x] = bits[x * ms + (ms - 1 - y)];
            bits = rot;
        }
    }
    return t;
}
static std::map<int, std::vector<uint8_t>> &table_cache() { static std::map<int, std::vector<uint8_t>> c; return c; }
}  // namespace

extern "C" int b2a_get_predefined_dictionary(int dict_id, b2a_dictionary *out)
{
    if (!out) return set_err(B2A_ERR_INVALID, "null output");
    for (const Predef &p : k_predef) {
        if (p.id != dict_id) continue;
        for (const Family &f : k_families) {
            if (std::strcmp(f.name, p.family)) continue;
            auto &cache = table_cache();
            auto it = cache.find(dict_id);
            if (it == cache.end()) it = cache.emplace(dict_id, build_table(f, p.nMarkers)).first;
            out->markerSize = f.markerSize; out->maxCorrectionBits = p.maxCorr; out->nMarkers = p.nMarkers;
            out->nBytes = f.nBytes; out->table = it->second.data();
            return B2A_OK;
        }
    }
    return set_err(B2A_ERR_INVALID, "unknown predefined dictionary id");
}

// ------------------------------------------------------------------------------------------------
// pose / observation kernels
// ------------------------------------------------------------------------------------------------
// 8 lanes per marker (one per row of the reprojection system), one marker per warp
constexpr int POSE_THREADS = 128;
struct Lanes8 {
    long long *marks = nullptr; mutable int nmark = 0;       // debug: clock64 trace of one marker
    __device__ __forceinline__ void mark() const { if (marks && (threadIdx.x & 7) == 0 && nmark < 30) marks[nmark++] = clock64(); }
    __device__ __forceinline__ int lane() const { return threadIdx.x & 7; }
    __device__ __forceinline__ int nlanes() const { return 8; }
    __device__ __forceinline__ void sync() const { __syncwarp(0xFFu << (threadIdx.x & 24)); }
};
__global__ void __launch_bounds__(POSE_THREADS)
k_pose(const float *__restrict__ corners, const int32_t *__restrict__ n_acc, int B, int max_markers,
       Camera cam, float marker_length, double *__restrict__ rvecs, double *__restrict__ tvecs, long long *marks)
{
    // one marker per WARP (its lanes 0-7): markers take different numbers of LM trials, and four markers in one warp would
    // execute each other's divergent paths one after the other
    __shared__ double s_sh[POSE_THREADS / 32][POSE_SH];
    const int total = B * max_markers;
    const int t = (int)((blockIdx.x * (unsigned)POSE_THREADS + threadIdx.x) >> 5);      // marker slot of this warp
    if (t >= total || (threadIdx.x & 31) >= 8) return;
    const int f = t / max_markers, m = t - f * max_markers;
    if (n_acc && m >= n_acc[f]) return;
    Lanes8 lg;
    if (marks && t == 0) { lg.marks = marks; lg.mark(); }
    solve_marker_pose(lg, cam, marker_length, corners + (size_t)t * 8, s_sh[threadIdx.x >> 5], rvecs + (size_t)t * 3, tvecs + (size_t)t * 3);
}

// Results go straight into the caller-visible pinned host arrays (device-accessible under unified
// addressing): one small launch that writes only the filled entries instead of eight capacity-sized copies.
struct ExportPtrs {
    const int32_t *n_acc, *n_rej, *status, *ids; const float *corners, *rejected; const double *rvecs, *tvecs;    // device
    int32_t *h_nacc, *h_nrej, *h_status, *h_ids; float *h_corners, *h_rejected; double *h_rvecs, *h_tvecs;        // pinned host
};
__global__ void __launch_bounds__(128)
k_export(ExportPtrs p, int max_markers, int with_pose)
{
    const int f = blockIdx.x, K = max_markers;
    const int na = p.n_acc[f], nr = p.n_rej[f];
    if (threadIdx.x == 0) { p.h_nacc[f] = na; p.h_nrej[f] = nr; p.h_status[f] = p.status[f]; }
    const size_t o = (size_t)f * K;
    for (int i = threadIdx.x; i < na; i += blockDim.x) p.h_ids[o + i] = p.ids[o + i];
    for (int i = threadIdx.x; i < na * 8; i += blockDim.x) p.h_corners[o * 8 + i] = p.corners[o * 8 + i];
    for (int i = threadIdx.x; i < nr * 8; i += blockDim.x) p.h_rejected[o * 8 + i] = p.rejected[o * 8 + i];
    if (with_pose)
        for (int i = threadIdx.x; i < na * 3; i += blockDim.x) { p.h_rvecs[o * 3 + i] = p.rvecs[o * 3 + i]; p.h_tvecs[o * 3 + i] = p.tvecs[o * 3 + i]; }
}

// n_dev (optional): the count lives on the device (the detector's n_accepted of the frame); the count is also written to *n_out
__global__ void k_observations(const float *__restrict__ corners, const int32_t *__restrict__ ids, const double *__restrict__ rvecs,
                               const double *__restrict__ tvecs, int n, const int32_t *__restrict__ n_dev, int cap, Camera cam, ObsParams op,
                               Observation *__restrict__ out, int *__restrict__ keep, int *__restrict__ n_out)
{
    if (n_dev) n = min(*n_dev, cap);
    if (n_out && blockIdx.x == 0 && threadIdx.x == 0) *n_out = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        keep[i] = make_observation(cam, op, corners + (size_t)i * 8, ids[i], rvecs + (size_t)i * 3, tvecs + (size_t)i * 3, out[i]) ? 1 : 0;
}

static Camera to_camera(const b2a_camera *c)
{
    Camera cam;
    cam.fx = c->K[0]; cam.fy = c->K[4]; cam.cx = c->K[2]; cam.cy = c->K[5];
    double D[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < c->nD && i < 5; ++i) D[i] = c->D[i];
    cam.k1 = D[0]; cam.k2 = D[1]; cam.p1 = D[2]; cam.p2 = D[3]; cam.k3 = D[4];
    return cam;
}

// ------------------------------------------------------------------------------------------------
// detector handle
// ------------------------------------------------------------------------------------------------
enum { ST_H2D, ST_GRAY, ST_THRESH, ST_ANCHORS, ST_SEGMENTS, ST_CYCLES, ST_SORT, ST_ASSIGN, ST_EMIT, ST_APPROX, ST_GROUP, ST_IDENT, ST_FINAL, ST_POSE, ST_D2H, ST_COUNT };
static const char *k_stage_names[ST_COUNT] = {"h2d", "bgr2gray", "threshold", "anchors", "segments", "cycles", "sort_scan", "assign", "emit",
                                              "approx", "group", "identify", "finalize", "pose", "d2h"};

struct b2a_detector {
    b2a_detector_config cfg;
    b2a_detector_params prm;
    b2a_dictionary dict;
    std::vector<uint8_t> dict_bytes;
    int device = 0, num_sms = 148;
    cudaStream_t stream = nullptr;            // == streams[0]
    static constexpr int MAX_SUB = 8;
    cudaStream_t streams[MAX_SUB] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[MAX_SUB] = {};
    int n_sub_max = MAX_SUB, n_streams = 0;       // 0 = automatic: 2 sub-batches for frames already in HBM, 4 when they still cross PCIe
    int nScales = 0, radius[MAX_SCALES];
    bool thresh_tiles = false;
    int tm_variant = 0;                           // marching threshold kernel: 0 = 24-row chunks (3 CTAs per SM), 1 / 2 = 12-row chunks (4 / 5 CTAs per SM)
    int max_cand = 0, max_markers = 0, surv_cap = 0;
    size_t gray_pitch = 0;
    // device memory
    uint8_t *d_in = nullptr, *d_gray = nullptr;
    int2 *d_rtab = nullptr;                       // ArUco3: the resize tables of each sub-batch ([MAX_SUB][max_width + max_height])
    uint8_t *d_pyr = nullptr, *d_segimg = nullptr;   // ArUco3: pyramid levels 1.. of every frame ([level][frame] planes) and the segmentation images
    uint32_t *d_masks = nullptr; size_t masks_words = 0;
    // border graph (core.h): anchors of all (frame,scale) masks of a sub-batch share one slice of these arrays
    uint2 *d_ast = nullptr; Seg *d_seg = nullptr, *d_sseg = nullptr; uint32_t *d_minoff = nullptr, *d_ssoff = nullptr; int2 *d_emit = nullptr; uint32_t *d_amap = nullptr;
    uint2 *d_starts = nullptr; uint32_t *d_codes = nullptr;
    unsigned anchors_cap = 0, starts_cap = 0; int anchor_R = 32;
    int *d_counters = nullptr;               // [sub-batch] n_anchors (unsigned), then per-(f,s) arrays
    unsigned *d_counters2 = nullptr;         // [sub-batch] n_starts
    int *d_surv_count = nullptr, *d_contour_count = nullptr, *d_iso_count = nullptr, *d_status = nullptr;
    uint4 *d_surv = nullptr, *d_sorted = nullptr;
    int *d_pts_off = nullptr; uint32_t *d_pts = nullptr; int pts_cap = 0;
    uint8_t *d_quad_ok = nullptr; int32_t *d_quad_xy = nullptr, *d_quad_len = nullptr;
    unsigned long long *d_dict = nullptr;
    double *d_wM = nullptr;                   // inverse perspective map of every identification work item
    unsigned long long *d_idcodes = nullptr;  // refineDetectedMarkers: extracted inner bits of one frame's work items
    WalkTables *d_tables = nullptr;
    FrameScratch fs0{};                       // frame-0 pointers
    FrameOutputs fo0{};
    float *d_corners2 = nullptr;              // corners after the optional refinement (subpix / contour lines)
    double *d_rvecs = nullptr, *d_tvecs = nullptr;
    // pinned host mirrors of the outputs
    int32_t *h_nacc = nullptr, *h_nrej = nullptr, *h_ids = nullptr, *h_status = nullptr;
    float *h_corners = nullptr, *h_rejected = nullptr;
    double *h_rvecs = nullptr, *h_tvecs = nullptr;
    // geometry of the last call (mask padding must be re-zeroed when it changes)
    int lastW = -1, lastH = -1, lastB = -1;
    // timing
    cudaEvent_t ev[ST_COUNT + 1];
    bool ev_used[ST_COUNT + 1];
    float stage_ms[ST_COUNT];
    int launches = 0;
    int timed_mode = 0;
    int last_call_batch = 0; bool last_call_pose = false;       // what the result arrays hold (b2a_detector_last_detections)
    // asynchronous submit / wait: the handle owns a second, lazily created pipeline context (all buffers and streams); batches
    // alternate between the two, so the H2D copy of one overlaps the kernels of the other
    static constexpr int MAX_CTX = 8;
    b2a_detector *more[MAX_CTX - 1] = {};         // contexts 1 .. n_ctx - 1 (this object is context 0)
    int n_ctx = 2;                                // batches in flight that submit / wait alternate over (b2a_detector_set_inflight)
    bool in_flight = false, pending_pose = false, pipelined = false; int pending_batch = 0;
    // small batches are launch-bound on the host (one 1080p frame: 19 launches, ~100 us of host time): the whole call is captured
    // once per (shape, camera) into a CUDA graph and replayed; the frames go through d_in so that every node reads fixed addresses
    struct GraphEntry {
        int B, W, H, channels, n_streams; bool has_cam, pipelined; b2a_camera cam;
        cudaGraphExec_t exec; int launches; unsigned long long last_use;
    };
    std::vector<GraphEntry> graphs;
    bool use_graph = true, graph_failed = false, capturing = false;
    unsigned long long graph_tick = 0;
    unsigned next_ticket = 0;
    std::vector<void *> allocs, pinned;
};

template <class T>
static int dev_alloc(b2a_detector *d, T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return set_err(B2A_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    d->allocs.push_back(q);
    *p = (T *)q;
    return B2A_OK;
}
template <class T>
static int pin_alloc(b2a_detector *d, T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMallocHost(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return set_err(B2A_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    d->pinned.push_back(q);
    *p = (T *)q;
    return B2A_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__ != B2A_OK) return rc__; } while (0)

extern "C" void b2a_detector_destroy(b2a_detector *d)
{
    if (!d) return;
    for (b2a_detector *&m : d->more) if (m) { b2a_detector_destroy(m); m = nullptr; }
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (void *p : d->allocs) cudaFree(p);
    for (void *p : d->pinned) cudaFreeHost(p);
    for (int i = 0; i <= ST_COUNT; ++i) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    for (int i = 0; i < b2a_detector::MAX_SUB; ++i) { if (d->streams[i]) cudaStreamDestroy(d->streams[i]); if (d->ev_join[i]) cudaEventDestroy(d->ev_join[i]); }
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    for (auto &g : d->graphs) cudaGraphExecDestroy(g.exec);
    delete d;
}

static int create_impl(b2a_detector *d)
{
    const b2a_detector_config &c = d->cfg;
    const b2a_detector_params &p = d->prm;
    CU(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c.device));
    d->num_sms = prop.multiProcessorCount;
    for (int i = 0; i < b2a_detector::MAX_SUB; ++i) {
        CU(cudaStreamCreateWithFlags(&d->streams[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&d->ev_join[i], cudaEventDisableTiming));
    }
    d->stream = d->streams[0];
    CU(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
    if (const char *e = std::getenv("B2A_STREAMS")) d->n_streams = std::max(1, std::min(std::atoi(e), b2a_detector::MAX_SUB));
    for (int i = 0; i <= ST_COUNT; ++i) CU(cudaEventCreate(&d->ev[i]));
    const int B = c.max_batch, W = c.max_width, H = c.max_height, nS = d->nScales;
    const size_t P = (size_t)W * H;
    d->gray_pitch = ((size_t)W + 15) & ~(size_t)15;
    TRY(dev_alloc(d, &d->d_in, (size_t)B * std::max(P * 3, (((size_t)W + 3) & ~(size_t)3) * H)));
    TRY(dev_alloc(d, &d->d_gray, (size_t)B * d->gray_pitch * H));
    if (p.useAruco3Detection) {
        Aruco3Plan plan;
        if (!aruco3_plan(W, H, p.minSideLengthCanonicalImg, p.minMarkerLengthRatioOriginalImg, plan))
            return set_err(B2A_ERR_UNSUPPORTED, "ArUco3: the image pyramid of the largest frame needs more than 12 levels, or the segmentation image is smaller than its last level");
        size_t per_frame = 0;
        for (int l = 1; l <= plan.numLevels; ++l) per_frame += (((size_t)plan.W[l] + 15) & ~(size_t)15) * plan.H[l];
        TRY(dev_alloc(d, &d->d_pyr, (size_t)B * per_frame));
        TRY(dev_alloc(d, &d->d_segimg, (size_t)B * d->gray_pitch * H));
        TRY(dev_alloc(d, &d->d_rtab, (size_t)b2a_detector::MAX_SUB * ((size_t)W + H)));
    }
    const int WW = (W + 31) / 32, PWW = WW + 2;
    d->masks_words = (size_t)B * nS * PWW * (H + 2);
    TRY(dev_alloc(d, &d->d_masks, d->masks_words));
    if (const char *e = std::getenv("B2A_ANCHOR_R")) { const int r = std::atoi(e); if (r >= 1 && r <= 32 && !(r & (r - 1))) d->anchor_R = r; }
    // anchors: a state is an anchor on every R-th row or column, so even a pure-noise mask (~0.8 states per pixel)
    // stays below 2 P / R per mask; start candidates: below P / 4 per mask.  Both arrays are cut per FRAME (run_front)
    d->anchors_cap = (unsigned)std::min<size_t>((size_t)B * std::max<size_t>(nS * (2 * P / (size_t)d->anchor_R), 1u << 16), 0x7FFFFFF0u);
    TRY(dev_alloc(d, &d->d_ast, d->anchors_cap)); TRY(dev_alloc(d, &d->d_seg, d->anchors_cap));
    TRY(dev_alloc(d, &d->d_minoff, d->anchors_cap)); TRY(dev_alloc(d, &d->d_emit, d->anchors_cap));
    TRY(dev_alloc(d, &d->d_sseg, d->anchors_cap)); TRY(dev_alloc(d, &d->d_ssoff, d->anchors_cap));
    TRY(dev_alloc(d, &d->d_amap, (size_t)B * nS * H * ((W + 31) / 32)));
    d->starts_cap = (unsigned)std::min<size_t>((size_t)B * std::max<size_t>(nS * (P / 4), 1u << 16), 0x7FFFFFF0u);
    TRY(dev_alloc(d, &d->d_starts, d->starts_cap));
    TRY(dev_alloc(d, &d->d_codes, (size_t)d->anchors_cap * SEG_CODE_WORDS));
    TRY(dev_alloc(d, &d->d_counters2, d->n_sub_max));
    const size_t FS = (size_t)B * nS;
    TRY(dev_alloc(d, &d->d_counters, d->n_sub_max + 3 * FS + B));
    d->d_surv_count = d->d_counters + d->n_sub_max; d->d_contour_count = d->d_surv_count + FS; d->d_iso_count = d->d_contour_count + FS;
    d->d_status = d->d_iso_count + FS;
    TRY(dev_alloc(d, &d->d_surv, FS * d->surv_cap));
    TRY(dev_alloc(d, &d->d_sorted, FS * d->surv_cap));
    TRY(dev_alloc(d, &d->d_pts_off, FS * d->surv_cap));
    d->pts_cap = (int)std::max<size_t>(P, 1 << 16);          // points of the kept borders of one mask: a speckled 1080p mask reaches 0.3 P
    TRY(dev_alloc(d, &d->d_pts, FS * (size_t)d->pts_cap));
    TRY(dev_alloc(d, &d->d_quad_ok, FS * d->surv_cap));
    TRY(dev_alloc(d, &d->d_quad_xy, FS * d->surv_cap * 8));
    TRY(dev_alloc(d, &d->d_quad_len, FS * d->surv_cap));
    // dictionary: one u64 per (marker, rotation), byte k in bits 8k..8k+7
    {
        std::vector<unsigned long long> packed((size_t)d->dict.nMarkers * 4, 0);
        for (int m = 0; m < d->dict.nMarkers; ++m) for (int r = 0; r < 4; ++r) {
            unsigned long long v = 0;
            for (int k = 0; k < d->dict.nBytes; ++k) v |= (unsigned long long)d->dict_bytes[((size_t)m * 4 + r) * d->dict.nBytes + k] << (8 * k);
            packed[(size_t)m * 4 + r] = v;
        }
        TRY(dev_alloc(d, &d->d_dict, packed.size()));
        CU(cudaMemcpy(d->d_dict, packed.data(), packed.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    {
        static WalkTables wt;
        for (int i = 0; i < 4096; ++i) build_walk_table_entry(wt, i);
        TRY(dev_alloc(d, &d->d_tables, 1));
        CU(cudaMemcpy(d->d_tables, &wt, sizeof(wt), cudaMemcpyHostToDevice));
    }
    const size_t MC = d->max_cand, BM = (size_t)B * MC;
    FrameScratch &fs = d->fs0;
    TRY(dev_alloc(d, &fs.cq, BM * 8)); TRY(dev_alloc(d, &fs.clen, BM)); TRY(dev_alloc(d, &fs.tq, BM * 8)); TRY(dev_alloc(d, &fs.tper, BM)); TRY(dev_alloc(d, &fs.cent, BM * 3));
    TRY(dev_alloc(d, &fs.gid, BM)); TRY(dev_alloc(d, &fs.sel, BM)); TRY(dev_alloc(d, &fs.gstart, (size_t)B * (MC + 1))); TRY(dev_alloc(d, &fs.gfill, BM));
    TRY(dev_alloc(d, &fs.members, BM)); TRY(dev_alloc(d, &fs.closeIdx, BM)); TRY(dev_alloc(d, &fs.closeCnt, BM));
    TRY(dev_alloc(d, &fs.S, BM)); TRY(dev_alloc(d, &fs.parent, BM)); TRY(dev_alloc(d, &fs.depth, BM)); TRY(dev_alloc(d, &fs.selGroup, BM));
    TRY(dev_alloc(d, &fs.closeM, BM * 2 * ((MC + 31) / 32)));       // M and its transpose
    TRY(dev_alloc(d, &fs.wq, BM * 8)); TRY(dev_alloc(d, &fs.wres, BM)); TRY(dev_alloc(d, &fs.closeStart, BM)); TRY(dev_alloc(d, &fs.closeNum, BM));
    TRY(dev_alloc(d, &fs.counters, (size_t)B * 8));
    if (p.useAruco3Detection) { TRY(dev_alloc(d, &fs.tlen, BM)); TRY(dev_alloc(d, &fs.wlen, BM)); }
    TRY(dev_alloc(d, &d->d_wM, BM * 9));
    TRY(dev_alloc(d, &d->d_idcodes, (size_t)d->max_cand));
    const size_t BK = (size_t)B * d->max_markers;
    FrameOutputs &fo = d->fo0;
    TRY(dev_alloc(d, &fo.n_accepted, B)); TRY(dev_alloc(d, &fo.n_rejected, B)); TRY(dev_alloc(d, &fo.status, B));
    TRY(dev_alloc(d, &fo.corners, BK * 8)); TRY(dev_alloc(d, &fo.ids, BK)); TRY(dev_alloc(d, &fo.rejected, BK * 8));
    TRY(dev_alloc(d, &d->d_corners2, BK * 8));
    TRY(dev_alloc(d, &d->d_rvecs, BK * 3)); TRY(dev_alloc(d, &d->d_tvecs, BK * 3));
    TRY(pin_alloc(d, &d->h_nacc, B)); TRY(pin_alloc(d, &d->h_nrej, B)); TRY(pin_alloc(d, &d->h_status, B));
    TRY(pin_alloc(d, &d->h_corners, BK * 8)); TRY(pin_alloc(d, &d->h_ids, BK)); TRY(pin_alloc(d, &d->h_rejected, BK * 8));
    TRY(pin_alloc(d, &d->h_rvecs, BK * 3)); TRY(pin_alloc(d, &d->h_tvecs, BK * 3));
    CU(cudaFuncSetAttribute(k_group, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CU(cudaFuncSetAttribute(k_group_a, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (d->max_cand + 1) * (int)sizeof(uint32_t)));
    CU(cudaFuncSetAttribute(k_identify<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)identify_smem_bytes(ID_MAX_S)));
    CU(cudaFuncSetAttribute(k_identify<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)identify_smem_bytes(ID_MAX_S)));
    CU(cudaFuncSetAttribute(k_threshold3<1, 6, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3_SMEM));
    CU(cudaFuncSetAttribute(k_threshold_march<1, 6, 11, 24, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmCfg<24>::SMEM));
    CU(cudaFuncSetAttribute(k_threshold_march<1, 6, 11, 12, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmCfg<12>::SMEM));
    CU(cudaFuncSetAttribute(k_threshold_march<1, 6, 11, 12, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmCfg<12>::SMEM));
    CU(cudaFuncSetAttribute(k_threshold_march<1, 6, 11, 12, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmCfg<12, true>::SMEM));
    if (const char *e = std::getenv("B2A_TM_VARIANT")) d->tm_variant = std::atoi(e);
    d->thresh_tiles = std::getenv("B2A_THRESH_TILES") != nullptr;          // A/B switch: the tiled kernel (k_threshold3) instead of the marching one
    CU(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (int)sizeof(int32_t) * d->max_cand));
    CU(cudaStreamSynchronize(d->stream));
    return B2A_OK;
}

extern "C" int b2a_detector_create(const b2a_detector_config *cfg, const b2a_dictionary *dict, const b2a_detector_params *params, b2a_detector **out)
{
    if (!cfg || !dict || !out || !dict->table) return set_err(B2A_ERR_INVALID, "null argument");
    if (cfg->max_width <= 0 || cfg->max_height <= 0 || cfg->max_batch <= 0) return set_err(B2A_ERR_INVALID, "bad geometry");
    if (cfg->max_width > 32767 || cfg->max_height > 32767) return set_err(B2A_ERR_UNSUPPORTED, "frame larger than 32767");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_err(B2A_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return set_err(B2A_ERR_INVALID, "bad device ordinal");
    b2a_detector_params prm;
    if (params) prm = *params; else b2a_default_detector_params(&prm);
    if (prm.markerBorderBits < 1) return set_err(B2A_ERR_INVALID, "markerBorderBits < 1");
    if (prm.cornerRefinementMethod < 0 || prm.cornerRefinementMethod > 3) return set_err(B2A_ERR_INVALID, "cornerRefinementMethod");
    if (prm.cornerRefinementMethod == 3) return set_err(B2A_ERR_UNSUPPORTED, "CORNER_REFINE_APRILTAG (the AprilTag quad detector is not part of this library)");
    if (prm.useAruco3Detection) {
        if (prm.minSideLengthCanonicalImg < 1 || !(prm.minMarkerLengthRatioOriginalImg >= 0.f)) return set_err(B2A_ERR_INVALID, "ArUco3 needs minSideLengthCanonicalImg >= 1 and minMarkerLengthRatioOriginalImg >= 0");
        prm.cornerRefinementMethod = 1;        // "always turn on corner refinement in case of Aruco3, due to upsampling" (detectMarkers)
    }
    if (prm.adaptiveThreshWinSizeMin < 3 || prm.adaptiveThreshWinSizeMax < prm.adaptiveThreshWinSizeMin || prm.adaptiveThreshWinSizeStep <= 0)
        return set_err(B2A_ERR_INVALID, "adaptiveThreshWinSize*");
    if (dict->markerSize < 1 || dict->markerSize * dict->markerSize > 64 || dict->nBytes != (dict->markerSize * dict->markerSize + 7) / 8)
        return set_err(B2A_ERR_UNSUPPORTED, "marker size (up to 8x8)");
    b2a_detector *d = new b2a_detector();
    std::memset(d->ev, 0, sizeof(d->ev));
    d->cfg = *cfg; d->prm = prm; d->device = cfg->device;
    d->dict = *dict;
    d->dict_bytes.assign(dict->table, dict->table + (size_t)dict->nMarkers * 4 * dict->nBytes);
    d->dict.table = d->dict_bytes.data();
    d->nScales = (prm.adaptiveThreshWinSizeMax - prm.adaptiveThreshWinSizeMin) / prm.adaptiveThreshWinSizeStep + 1;
    if (d->nScales > MAX_SCALES) { delete d; return set_err(B2A_ERR_UNSUPPORTED, "more than 8 threshold scales"); }
    for (int i = 0; i < d->nScales; ++i) {
        int k = prm.adaptiveThreshWinSizeMin + i * prm.adaptiveThreshWinSizeStep;
        if (k % 2 == 0) k++;                                   // OpenCV bumps even window sizes to the next odd
        d->radius[i] = k / 2;
        if (d->radius[i] > R_MAX) { delete d; return set_err(B2A_ERR_UNSUPPORTED, "threshold window larger than 31"); }
    }
    const int S = (dict->markerSize + 2 * prm.markerBorderBits) * prm.perspectiveRemovePixelPerCell;
    if (S > ID_MAX_S || dict->markerSize + 2 * prm.markerBorderBits > 9) { delete d; return set_err(B2A_ERR_UNSUPPORTED, "warped marker image larger than 72 px"); }
    d->max_markers = cfg->max_markers > 0 ? cfg->max_markers : 256;
    d->max_cand = cfg->max_candidates > 0 ? cfg->max_candidates : 2048;
    d->max_cand = (d->max_cand + 31) & ~31;
    if (d->max_cand > 4096) { delete d; return set_err(B2A_ERR_UNSUPPORTED, "max_candidates > 4096"); }
    d->surv_cap = SORT_CAP;
    int rc = create_impl(d);
    if (rc != B2A_OK) { std::string keep = g_err; b2a_detector_destroy(d); g_err = keep; return rc; }
    *out = d;
    return B2A_OK;
}

extern "C" int b2a_detector_num_scales(const b2a_detector *d) { return d ? d->nScales : 0; }
extern "C" void *b2a_detector_stream(const b2a_detector *d) { return d ? (void *)d->stream : nullptr; }
extern "C" int b2a_last_launch_count(const b2a_detector *d) { return d ? d->launches : 0; }
extern "C" int b2a_last_stage_times(const b2a_detector *d, const char **names, float *ms, int cap)
{
    if (!d) return 0;
    int n = 0;
    for (int i = 0; i < ST_COUNT && n < cap; ++i) { if (names) names[n] = k_stage_names[i]; if (ms) ms[n] = d->stage_ms[i]; ++n; }
    return n;
}

// ------------------------------------------------------------------------------------------------
// the pipeline.  A call's batch is cut into sub-batches that run on separate CUDA streams: the
// contour walks, the per-frame grouping, the identification and the pose chains are latency-bound
// (few long dependent chains), so sub-batches overlap their tails with each other's bulk work.
// Frames are independent; every per-frame / per-(frame,scale) array is simply addressed from the
// sub-batch's first frame, and each sub-batch owns a slice of the start-candidate list.
// ------------------------------------------------------------------------------------------------
struct Sub {
    int sb;                  // sub-batch index
    int b0, nb;              // first frame, number of frames
    cudaStream_t st;
    bool timed;              // stage events are recorded for sub-batch 0 only
    cudaEvent_t tl_after_h2d = nullptr;   // debug timeline
    DetGeom g;               // geometry with B = nb
    const uint8_t *gray; size_t pitch, frame_stride;      // gray frames of this sub-batch (ArUco3: the segmentation images)
    PyrLevels pyr;           // ArUco3: the pyramid of the full-size gray frames (n = 0: off)
    int closestIdx = 0;      // ArUco3: the pyramid level closest to the segmentation image
};

static int check_frames(b2a_detector *d, const b2a_frames *f)
{
    if (!d || !f || !f->data) return set_err(B2A_ERR_INVALID, "null argument");
    if (f->batch <= 0 || f->batch > d->cfg.max_batch) return set_err(B2A_ERR_INVALID, "batch outside [1, max_batch]");
    if (f->width <= 0 || f->height <= 0 || f->width > d->cfg.max_width || f->height > d->cfg.max_height || (size_t)f->width * f->height > (size_t)d->cfg.max_width * d->cfg.max_height)
        return set_err(B2A_ERR_INVALID, "frame size outside the handle's maximum");
    if (f->channels != 1 && f->channels != 3) return set_err(B2A_ERR_INVALID, "channels must be 1 or 3");
    if (f->row_stride && f->row_stride < (size_t)f->width * f->channels) return set_err(B2A_ERR_INVALID, "row_stride too small");
    return B2A_OK;
}

static void stage_mark(b2a_detector *d, const Sub &s, int st)
{
    if (!s.timed) return;
    cudaEventRecord(d->ev[st], s.st);
    d->ev_used[st] = true;
}

// B2A_SYNC_DEBUG=1: synchronize after every front-end kernel so that a device fault names its kernel
static bool sync_debug() { static const bool on = std::getenv("B2A_SYNC_DEBUG") != nullptr; return on; }
#define DBG_SYNC(st)                                                                                           \
    do {                                                                                                       \
        if (sync_debug()) {                                                                                    \
            cudaError_t e__ = cudaStreamSynchronize(st);                                                       \
            if (e__ != cudaSuccess) return set_err(B2A_ERR_CUDA, std::string("after launch ") + std::to_string(d->launches) + \
                                                   " of the front end (line " + std::to_string(__LINE__) + "): " + cudaGetErrorString(e__)); \
        }                                                                                                      \
    } while (0)

static int launch_err(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(B2A_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return B2A_OK;
}

static DetGeom make_geom(const b2a_detector *d, int W, int H, int B)
{
    DetGeom g;
    g.W = W; g.H = H; g.B = B; g.nScales = d->nScales;
    for (int i = 0; i < MAX_SCALES; ++i) g.radius[i] = i < d->nScales ? d->radius[i] : 0;
    g.Cfloor = (int)std::floor(d->prm.adaptiveThreshConstant);
    g.WW = (W + 31) / 32; g.PWW = g.WW + 2; g.mask_plane = (long long)g.PWW * (H + 2);
    g.KS = W + 1;
    g.maxWH = std::max(W, H);
    g.minPerim = (int)(unsigned)(d->prm.minMarkerPerimeterRate * g.maxWH);
    g.maxPerim = (int)(unsigned)(d->prm.maxMarkerPerimeterRate * g.maxWH);
    g.approxRate = d->prm.polygonalApproxAccuracyRate; g.minCornerDistRate = d->prm.minCornerDistanceRate;
    g.surv_cap = d->surv_cap; g.pts_cap = d->pts_cap; g.count_all = 0;
    return g;
}

static FrameScratch offset_scratch(const FrameScratch &b, size_t f, size_t mc)
{
    FrameScratch s = b;
    const size_t o = f * mc;
    s.cq += o * 8; s.clen += o; s.tq += o * 8; s.tper += o; s.cent += o * 3; s.gid += o; s.sel += o;
    s.gstart += f * (mc + 1); s.gfill += o; s.members += o; s.closeIdx += o; s.closeCnt += o;
    s.S += o; s.parent += o; s.depth += o; s.selGroup += o;
    s.closeM += o * 2 * ((mc + 31) / 32);
    s.wq += o * 8; s.wres += o; s.closeStart += o; s.closeNum += o;
    s.counters += f * 8;
    if (s.tlen) { s.tlen += o; s.wlen += o; }
    return s;
}

// ingest + A1 + A2 + A3 for one sub-batch: everything up to the quads of every (frame, scale)
static int run_front(b2a_detector *d, const b2a_frames *f, Sub &s, int walk_max_len /* 0 = maxPerimeter */)
{
    int W = f->width, H = f->height;
    const int nb = s.nb, b0 = s.b0;
    const bool a3 = d->prm.useAruco3Detection != 0;
    const size_t in_pitch = f->row_stride ? f->row_stride : (size_t)W * f->channels;
    const size_t in_frame = f->frame_stride ? f->frame_stride : in_pitch * H;
    cudaStream_t st = s.st;
    stage_mark(d, s, ST_H2D);
    const uint8_t *src = f->data + (size_t)b0 * in_frame;
    size_t src_pitch = in_pitch, src_frame = in_frame;
    if (!f->on_device) {
        const size_t rowbytes = (size_t)W * f->channels;
        const size_t dpitch = f->channels == 1 ? ((rowbytes + 3) & ~(size_t)3) : rowbytes;        // gray rows land 4-byte aligned (the marching threshold kernel loads words)
        uint8_t *dst = d->d_in + (size_t)b0 * dpitch * H;
        if (in_frame == in_pitch * H && in_pitch == rowbytes && dpitch == rowbytes) CU(cudaMemcpyAsync(dst, src, rowbytes * H * nb, cudaMemcpyHostToDevice, st));      // contiguous frames: one linear copy
        else if (in_frame == in_pitch * H) CU(cudaMemcpy2DAsync(dst, dpitch, src, in_pitch, rowbytes, (size_t)H * nb, cudaMemcpyHostToDevice, st));
        else for (int b = 0; b < nb; ++b) CU(cudaMemcpy2DAsync(dst + (size_t)b * dpitch * H, dpitch, src + (size_t)b * in_frame, in_pitch, rowbytes, H, cudaMemcpyHostToDevice, st));
        src = dst; src_pitch = dpitch; src_frame = dpitch * H;
    }
    if (s.tl_after_h2d) cudaEventRecord(s.tl_after_h2d, st);
    stage_mark(d, s, ST_GRAY);
    s.g = make_geom(d, W, H, nb);
    DetGeom &g = s.g;
    const bool default_windows = g.nScales == 3 && g.radius[0] == 1 && g.radius[1] == 6 && g.radius[2] == 11 && std::abs(g.Cfloor) <= 2048;
    // bgr8 frames whose rows are whole words go straight into the marching threshold kernel, which converts on the fly and leaves the
    // gray plane for the later stages (S0 fused into S1); anything else takes the separate conversion pass
    static const bool fuse_env = !(std::getenv("B2A_FUSE_BGR") && std::atoi(std::getenv("B2A_FUSE_BGR")) == 0);
    const bool fuse_bgr = f->channels == 3 && fuse_env && default_windows && !d->thresh_tiles && (W & 3) == 0 && (src_pitch & 3) == 0 && (src_frame & 3) == 0 &&
                          (((size_t)src) & 3) == 0 && src_pitch < (1ull << 32) && !a3;     // ArUco3 needs the gray plane before the threshold
    if (f->channels == 3) {
        uint8_t *gdst = d->d_gray + (size_t)b0 * d->gray_pitch * H;
        if (!fuse_bgr) {
            k_bgr2gray<<<d->num_sms * 4, 256, 0, st>>>(src, src_pitch, src_frame, gdst, d->gray_pitch, d->gray_pitch * H, W, H, nb);
            d->launches++;
        }
        s.gray = gdst; s.pitch = d->gray_pitch; s.frame_stride = d->gray_pitch * H;
    } else { s.gray = src; s.pitch = src_pitch; s.frame_stride = src_frame; }
    s.pyr.n = 0;
    if (a3) {
        // the pyramid of the full-size gray frames (levels 1.. live in d_pyr as [level][frame] planes) and, when the factor is
        // not 1, the reduced segmentation image that every later stage of the front end works on
        Aruco3Plan plan;
        if (!aruco3_plan(W, H, d->prm.minSideLengthCanonicalImg, d->prm.minMarkerLengthRatioOriginalImg, plan))
            return set_err(B2A_ERR_INVALID, "ArUco3: for this frame size the segmentation image is smaller than the last pyramid level (cv2 reads past its pyramid there)");
        PyrLevels &pl = s.pyr;
        pl.n = plan.numLevels + 1; pl.segW = plan.segW; pl.minPerimeter = 4 * d->prm.minSideLengthCanonicalImg;
        pl.W[0] = W; pl.H[0] = H; pl.base[0] = s.gray; pl.pitch[0] = s.pitch; pl.frame_stride[0] = s.frame_stride;
        s.closestIdx = plan.closestIdx;
        size_t off = 0;
        for (int l = 1; l <= plan.numLevels; ++l) {
            const size_t lp = ((size_t)plan.W[l] + 15) & ~(size_t)15, lf = lp * plan.H[l];
            pl.W[l] = plan.W[l]; pl.H[l] = plan.H[l]; pl.base[l] = d->d_pyr + off + (size_t)b0 * lf; pl.pitch[l] = lp; pl.frame_stride[l] = lf;
            off += (size_t)d->cfg.max_batch * lf;
        }
        // the large levels by the whole grid, one launch each; once a frame's level is a few thousand 4-pixel groups, the rest of the
        // pyramid by one CTA per frame in one launch (measured: a CTA per frame from level 2 on cost 0.13 ms more per 32 frames)
        int l = 1;
        for (; l <= plan.numLevels && ((plan.W[l] + 3) / 4) * plan.H[l] > 4096; ++l) {
            k_pyr_down<<<(unsigned)std::min<long long>(((long long)nb * plan.H[l] + 3) / 4, (long long)d->num_sms * 8), dim3(64, 4), 0, st>>>(pl.base[l - 1], pl.W[l - 1], pl.H[l - 1], pl.pitch[l - 1], pl.frame_stride[l - 1],
                                                                                                                   const_cast<uint8_t *>(pl.base[l]), pl.pitch[l], pl.frame_stride[l], nb);
            d->launches++;
        }
        if (l <= plan.numLevels) { k_pyr_chain<<<nb, 1024, 0, st>>>(pl, l); d->launches++; }
        if (plan.fxfy != 1.f) {
            const size_t sp = ((size_t)plan.segW + 15) & ~(size_t)15, sf = sp * plan.segH;
            uint8_t *seg = d->d_segimg + (size_t)b0 * sf;
            int2 *tab = d->d_rtab + (size_t)s.sb * ((size_t)d->cfg.max_width + d->cfg.max_height);
            k_resize_tabs<<<(plan.segW + plan.segH + 255) / 256, 256, 0, st>>>(W, H, plan.segW, plan.segH, tab);
            k_resize_linear<<<(unsigned)std::min<long long>(((long long)nb * plan.segH + 3) / 4, (long long)d->num_sms * 8), dim3(64, 4), 0, st>>>(s.gray, W, H, s.pitch, s.frame_stride, seg, plan.segW, plan.segH, sp, sf, nb, tab);
            d->launches += 2;
            s.gray = seg; s.pitch = sp; s.frame_stride = sf;
            W = plan.segW; H = plan.segH;
        }
        s.g = make_geom(d, W, H, nb);
        g.minPerim = pl.minPerimeter;          // _findMarkerContours: "for aruco3 we want to filter contours with min size"
        DBG_SYNC(st);
    }
    g.count_all = walk_max_len > 0 ? 1 : 0;          // the contour tap (exact counts, no give-up length) is the only caller that asks
    // this sub-batch's slice of the anchor arrays, and its per-(frame,scale) arrays addressed from frame b0
    const size_t fs0 = (size_t)b0 * g.nScales, FS = (size_t)nb * g.nScales;
    // a sub-batch owns the share of the anchor / start-candidate arrays that belongs to its frames: the arrays are sized per frame
    // (2P/R anchors and P/4 start candidates per mask), so any cut of the batch leaves every frame its full capacity
    const unsigned a_per = d->anchors_cap / (unsigned)d->cfg.max_batch, s_per = d->starts_cap / (unsigned)d->cfg.max_batch;
    const size_t a_off = (size_t)b0 * a_per, s_off = (size_t)b0 * s_per;
    const unsigned slice = a_per * (unsigned)nb;
    BorderGraph bg;
    bg.ast = d->d_ast + a_off; bg.seg = d->d_seg + a_off; bg.minoff = d->d_minoff + a_off;
    bg.sseg = d->d_sseg + a_off; bg.ssoff = d->d_ssoff + a_off;
    bg.emit = d->d_emit + a_off; bg.amap = d->d_amap + fs0 * (size_t)H * g.WW;
    bg.n_anchors = (unsigned *)d->d_counters + s.sb; bg.cap = slice;
    bg.codes = d->d_codes + a_off * SEG_CODE_WORDS;
    bg.starts = d->d_starts + s_off; bg.n_starts = d->d_counters2 + s.sb;
    bg.starts_cap = s_per * (unsigned)nb;
    const int Rm = d->anchor_R - 1, Rm2 = 8 * d->anchor_R - 1, max_len = walk_max_len > 0 ? walk_max_len : g.maxPerim;
    static const int walk_ctas = std::getenv("B2A_WALK_CTAS") ? std::max(1, std::atoi(std::getenv("B2A_WALK_CTAS"))) : 8;   // resident 256-thread CTAs per SM of the grid-stride kernels
    const unsigned walk_grid = (unsigned)d->num_sms * walk_ctas;
    uint32_t *masks = d->d_masks + fs0 * g.mask_plane;
    stage_mark(d, s, ST_THRESH);
    {
        const bool aligned4 = (s.pitch & 3) == 0 && (s.frame_stride & 3) == 0 && (((size_t)s.gray) & 3) == 0 && s.pitch < (1ull << 32);
        if (fuse_bgr) {
            // same work items as below with 12-row chunks (the staged rows are three times as wide) at 4 CTAs per SM
            const int n_sx = (W + TM_WT - 1) / TM_WT, slots = d->num_sms * 4;
            int Hs = TM_RC;
            double best = -1.0;
            for (int h = TM_RC; h <= TM_MAX_HS; h += TM_RC) {
                const int n = nb * n_sx * ((H + h - 1) / h);
                const double score = (double)n / ((double)((n + slots - 1) / slots) * slots) / (1.0 + 0.3 * 22.0 / h);
                if (score > best * 1.0001) { best = score; Hs = h; }
            }
            const int n_sy = (H + Hs - 1) / Hs, n_items = nb * n_sx * n_sy;
            k_threshold_march<1, 6, 11, 12, 4, true><<<std::min(n_items, slots), TM_THREADS, TmCfg<12, true>::SMEM, st>>>(
                src, (uint32_t)src_pitch, src_frame, masks, g, Hs, n_sy, n_sx, n_items, const_cast<uint8_t *>(s.gray), (uint32_t)s.pitch, s.frame_stride);
        } else if (default_windows && aligned4 && !d->thresh_tiles) {
            // marching kernel: work items = (frame, 320-column strip, Hs-row segment); Hs is the tallest segment that still gives every
            // resident CTA slot an item (a segment costs 22 rows of prefix warm-up)
            const int ctas_per_sm = d->tm_variant == 0 ? 3 : d->tm_variant == 1 ? 4 : 5;
            const int n_sx = (W + TM_WT - 1) / TM_WT, slots = d->num_sms * ctas_per_sm;
            int Hs = TM_RC;
            double best = -1.0;
            for (int h = TM_RC; h <= TM_MAX_HS; h += TM_RC) {
                // share of the resident slots that is busy over the whole launch (items run in waves) x share of rows that are not warm-up
                const int n = nb * n_sx * ((H + h - 1) / h);
                const double score = (double)n / ((double)((n + slots - 1) / slots) * slots) / (1.0 + 0.3 * 22.0 / h);
                if (score > best * 1.0001) { best = score; Hs = h; }
            }
            const int n_sy = (H + Hs - 1) / Hs, n_items = nb * n_sx * n_sy;
            if (d->tm_variant == 0) k_threshold_march<1, 6, 11, 24, 3><<<std::min(n_items, slots), TM_THREADS, TmCfg<24>::SMEM, st>>>(s.gray, (uint32_t)s.pitch, s.frame_stride, masks, g, Hs, n_sy, n_sx, n_items);
            else if (d->tm_variant == 1) k_threshold_march<1, 6, 11, 12, 4><<<std::min(n_items, slots), TM_THREADS, TmCfg<12>::SMEM, st>>>(s.gray, (uint32_t)s.pitch, s.frame_stride, masks, g, Hs, n_sy, n_sx, n_items);
            else k_threshold_march<1, 6, 11, 12, 5><<<std::min(n_items, slots), TM_THREADS, TmCfg<12>::SMEM, st>>>(s.gray, (uint32_t)s.pitch, s.frame_stride, masks, g, Hs, n_sy, n_sx, n_items);
        } else if (default_windows) {
            dim3 grid((W + T3_TW - 1) / T3_TW, (H + T3_TH - 1) / T3_TH, nb);
            k_threshold3<1, 6, 11><<<grid, T3_THREADS, T3_SMEM, st>>>(s.gray, s.pitch, s.frame_stride, masks, g);
        } else {
            dim3 grid((W + TH_TW - 1) / TH_TW, (H + TH_TH - 1) / TH_TH, nb);
            k_threshold<<<grid, TH_THREADS, 0, st>>>(s.gray, s.pitch, s.frame_stride, masks, g);
        }
        d->launches++;
    }
    stage_mark(d, s, ST_ANCHORS);
    k_anchors<<<walk_grid, 256, 0, st>>>(masks, bg, d->d_iso_count + fs0, Rm, Rm2, g);
    d->launches++; DBG_SYNC(st);
    stage_mark(d, s, ST_SEGMENTS);
    k_segments<<<walk_grid, 256, 0, st>>>(masks, bg, max_len, d->d_tables, Rm, g);
    d->launches++; DBG_SYNC(st);
    stage_mark(d, s, ST_CYCLES);
    k_skip<<<walk_grid, 256, 0, st>>>(bg, max_len);
    d->launches++; DBG_SYNC(st);
    k_cycles<<<walk_grid, 256, 0, st>>>(masks, bg, d->d_surv + fs0 * g.surv_cap, d->d_surv_count + fs0, d->d_contour_count + fs0, max_len, d->d_tables, Rm, g);
    d->launches++; DBG_SYNC(st);
    stage_mark(d, s, ST_SORT);
    k_sort_scan<<<(unsigned)FS, 1024, 0, st>>>(d->d_surv + fs0 * g.surv_cap, d->d_surv_count + fs0, d->d_sorted + fs0 * g.surv_cap,
                                               d->d_pts_off + fs0 * g.surv_cap, d->d_status + b0, bg.n_anchors, bg.cap, bg.n_starts, bg.starts_cap, g);
    d->launches++; DBG_SYNC(st);
    stage_mark(d, s, ST_ASSIGN);
    static const int assign_ctas = std::getenv("B2A_ASSIGN_CTAS") ? std::max(1, std::atoi(std::getenv("B2A_ASSIGN_CTAS"))) : 8;
    k_assign<<<dim3(assign_ctas, (unsigned)FS), 128, 0, st>>>(masks, bg, d->d_sorted + fs0 * g.surv_cap, d->d_surv_count + fs0, d->d_pts_off + fs0 * g.surv_cap,
                                                    d->d_pts + fs0 * (size_t)g.pts_cap, d->d_tables, g);
    d->launches++; DBG_SYNC(st);
    k_assign_sub<<<walk_grid, 256, 0, st>>>(bg);
    d->launches++; DBG_SYNC(st);
    stage_mark(d, s, ST_EMIT);
    k_emit<<<walk_grid, 256, 0, st>>>(masks, bg, d->d_sorted + fs0 * g.surv_cap, d->d_pts_off + fs0 * g.surv_cap, d->d_pts + fs0 * (size_t)g.pts_cap, d->d_tables, g);
    d->launches++; DBG_SYNC(st);
    stage_mark(d, s, ST_APPROX);
    static const int approx_short = std::getenv("B2A_APPROX_SHORT") ? std::max(1, std::atoi(std::getenv("B2A_APPROX_SHORT"))) : 32;   // CTAs (8 warps each) per (frame,scale) for the warp-per-border role: a (frame,scale) has 200-350 kept borders of very different lengths, and with 64 warps the longest queue set the kernel's time (8 -> 32 CTAs: 0.185 -> 0.156 ms)
    k_approx<<<dim3(APPROX_LONG_CTAS + approx_short, (unsigned)FS), 256, 0, st>>>(d->d_sorted + fs0 * g.surv_cap, d->d_surv_count + fs0, d->d_pts_off + fs0 * g.surv_cap,
                                                                       d->d_pts + fs0 * (size_t)g.pts_cap, d->d_quad_ok + fs0 * g.surv_cap, d->d_quad_xy + fs0 * g.surv_cap * 8,
                                                                       d->d_quad_len + fs0 * g.surv_cap, g);
    d->launches++; DBG_SYNC(st);
    return launch_err("front-end kernels");
}

static FrameParams frame_params(const b2a_detector *d, const DetGeom &g)
{
    FrameParams fp;
    fp.W = g.W; fp.H = g.H; fp.nScales = g.nScales; fp.surv_cap = g.surv_cap; fp.max_cand = d->max_cand; fp.max_markers = d->max_markers;
    fp.markerSize = d->dict.markerSize; fp.borderBits = d->prm.markerBorderBits; fp.minDistanceToBorder = d->prm.minDistanceToBorder;
    fp.minMarkerDistanceRate = (float)d->prm.minMarkerDistanceRate; fp.minGroupDistance = d->prm.minGroupDistance;
    fp.detectInverted = d->prm.detectInvertedMarker ? 1 : 0;
    return fp;
}

// cornerSubPix of the accepted corners: a warp per corner while the largest window's patch fits the default shared memory
// (windows up to 25), else a thread per corner
static void launch_subpix(b2a_detector *d, cudaStream_t st, const uint8_t *gray, const int32_t *n_acc, const float *cin, float *cout, int nb, const SubpixParams &sp)
{
    const int wmax = sp.fixedWin > 0 ? sp.fixedWin : std::max(1, sp.maxWin);
    const size_t smem = subpix_warp_smem_bytes(wmax);
    static const bool thread_form = std::getenv("B2A_SUBPIX_THREAD") != nullptr;          // A/B switch
    if (smem <= 48 * 1024 && !thread_form) k_subpix_warp<<<d->num_sms * 8, 128, smem, st>>>(gray, n_acc, cin, cout, nb, sp, wmax);
    else k_subpix<<<d->num_sms * 2, 128, 0, st>>>(gray, n_acc, cin, cout, nb, sp);
    d->launches++;
}

static int run_back(b2a_detector *d, Sub &s, const b2a_camera *cam, bool stop_after_group)
{
    const DetGeom &g = s.g;
    cudaStream_t st = s.st;
    const int b0 = s.b0, nb = s.nb;
    const size_t fs0 = (size_t)b0 * g.nScales, K = d->max_markers;
    FrameArrays fa;
    fa.fs0 = offset_scratch(d->fs0, b0, d->max_cand);
    fa.fo0 = d->fo0;
    fa.fo0.n_accepted += b0; fa.fo0.n_rejected += b0; fa.fo0.corners += (size_t)b0 * K * 8; fa.fo0.ids += (size_t)b0 * K; fa.fo0.rejected += (size_t)b0 * K * 8;
    fa.fo0.status = d->d_status + b0;
    fa.surv_count = d->d_surv_count + fs0; fa.quad_ok = d->d_quad_ok + fs0 * g.surv_cap; fa.quad_xy = d->d_quad_xy + fs0 * g.surv_cap * 8;
    fa.quad_len = d->d_quad_len + fs0 * g.surv_cap;
    fa.marks = nullptr;
#ifdef B2A_DEBUG_TAPS
    static long long *dbg_marks = nullptr;
    if (std::getenv("B2A_GROUP_MARKS")) {
        if (!dbg_marks) { cudaMalloc(&dbg_marks, (size_t)d->cfg.max_batch * 32 * sizeof(long long)); }
        cudaMemsetAsync(dbg_marks, 0, (size_t)d->cfg.max_batch * 32 * sizeof(long long), st);
        fa.marks = dbg_marks + (size_t)b0 * 32;
    }
#endif
    const FrameParams fp = frame_params(d, g);
    stage_mark(d, s, ST_GROUP);
    const int smem_words = 50 * 1024;                       // 200 KB: per-candidate arrays + closeness matrix
    k_group_a<<<nb, 1024, 8 * (d->max_cand + 1) * sizeof(uint32_t), st>>>(fa, fp);
    k_close<<<dim3(CLOSE_CTAS, nb), 256, 0, st>>>(fa, fp);
    k_group<<<nb, 1024, smem_words * sizeof(uint32_t), st>>>(fa, fp, smem_words);
    d->launches += 3;
#ifdef B2A_DEBUG_TAPS
    if (fa.marks && s.sb == 0) {
        long long hm[32];
        cudaStreamSynchronize(st);
        cudaMemcpy(hm, fa.marks, sizeof(hm), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "k_group frame %d sync-to-sync cycles:", b0);
        for (int i = 1; i < 32 && hm[i]; ++i) std::fprintf(stderr, hm[i] < 0 ? " (%lld)" : " %lld", std::llabs(hm[i]) - std::llabs(hm[i - 1]));
        std::fprintf(stderr, "\n");
    }
#endif
    if (stop_after_group) return launch_err("k_group");
    stage_mark(d, s, ST_IDENT);
    IdentParams ip;
    ip.markerSize = d->dict.markerSize; ip.borderBits = d->prm.markerBorderBits; ip.cellSize = d->prm.perspectiveRemovePixelPerCell;
    ip.cellMargin = (int)(d->prm.perspectiveRemoveIgnoredMarginPerCell * ip.cellSize);
    ip.nMarkers = d->dict.nMarkers; ip.maxCorr = (int)((double)d->dict.maxCorrectionBits * d->prm.errorCorrectionRate);
    ip.maxBorderErr = (int)(d->dict.markerSize * d->dict.markerSize * d->prm.maxErroneousBitsInBorderRate);
    ip.detectInverted = d->prm.detectInvertedMarker ? 1 : 0;
    ip.minOtsuStdDev = d->prm.minOtsuStdDev; ip.W = g.W; ip.H = g.H; ip.pitch = s.pitch; ip.frame_stride = s.frame_stride; ip.max_cand = d->max_cand;
    ip.marks = nullptr; ip.codes = nullptr;
    ip.pyr = s.pyr;
#ifdef B2A_DEBUG_TAPS
    static long long *id_marks = nullptr;
    if (std::getenv("B2A_IDENT_MARKS") && s.sb == 0) {
        const size_t nrec = 16 + (size_t)d->cfg.max_batch * 128 * 3;
        if (!id_marks) cudaMalloc(&id_marks, nrec * sizeof(long long));
        cudaMemsetAsync(id_marks, 0, nrec * sizeof(long long), st);
        ip.marks = id_marks;
    }
#endif
    double *wM = d->d_wM + (size_t)b0 * d->max_cand * 9;
    k_homography<<<dim3(4, nb), 64, 0, st>>>(fa, wM, (ip.markerSize + 2 * ip.borderBits) * ip.cellSize, d->max_cand, s.pyr);
    d->launches++;
    static const int id_blocks = std::getenv("B2A_ID_BLOCKS") ? std::max(1, std::atoi(std::getenv("B2A_ID_BLOCKS"))) : 48;   // x 4 warps = work items of a frame in flight (a frame has ~124 of them; with 128 warps the frames that have more made a second round: 0.118 -> 0.102 ms)
    if (s.pyr.n > 0) k_identify<true><<<dim3(id_blocks, nb), ID_THREADS, identify_smem_bytes((ip.markerSize + 2 * ip.borderBits) * ip.cellSize), st>>>(s.gray, d->d_dict, wM, fa, ip);
    else k_identify<false><<<dim3(id_blocks, nb), ID_THREADS, identify_smem_bytes((ip.markerSize + 2 * ip.borderBits) * ip.cellSize), st>>>(s.gray, d->d_dict, wM, fa, ip);
    d->launches++;
#ifdef B2A_DEBUG_TAPS
    if (ip.marks) {
        long long hm[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(hm, ip.marks, sizeof(hm), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "k_identify item 0 phase cycles (sample, hist, otsu-prep, chain, sigma, rest):");
        for (int i = 1; i < 16 && hm[i]; ++i) std::fprintf(stderr, " %lld", hm[i] - hm[i - 1]);
        std::fprintf(stderr, "\n");
        std::vector<long long> rec((size_t)nb * 128 * 3);
        cudaMemcpy(rec.data(), ip.marks + 16, rec.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        long long t0 = 0, t1 = 0; int cnt = 0; std::vector<long long> dur;
        for (size_t i = 0; i < rec.size(); i += 3) if (rec[i]) { if (!t0 || rec[i] < t0) t0 = rec[i]; if (rec[i + 1] > t1) t1 = rec[i + 1]; dur.push_back(rec[i + 1] - rec[i]); ++cnt; }
        std::sort(dur.begin(), dur.end());
        long long latest_start = 0;
        for (size_t i = 0; i < rec.size(); i += 3) if (rec[i] && rec[i] - t0 > latest_start) latest_start = rec[i] - t0;
        if (cnt) std::fprintf(stderr, "k_identify items %d: span %lld ns, item ns min %lld median %lld p90 %lld max %lld, latest start +%lld ns\n", cnt, t1 - t0, dur[0], dur[cnt / 2],
                              dur[cnt * 9 / 10], dur[cnt - 1], latest_start);
    }
#endif
    stage_mark(d, s, ST_FINAL);
    k_finalize<<<nb, 128, 8 * sizeof(int32_t) * d->max_cand, st>>>(fa, fp);
    d->launches++;
    float *corners = fa.fo0.corners;
    if (s.pyr.n > 0) {
        // ArUco3, findCornerInPyrImage: the accepted corners go to the pyramid level closest to the segmentation image and are then
        // doubled and refined level by level down to the full-size frame; with closestIdx = 0 they are only scaled
        const PyrLevels &pl = s.pyr;
        SubpixParams sp;
        sp.max_markers = d->max_markers; sp.markerSize = d->dict.markerSize; sp.borderBits = d->prm.markerBorderBits; sp.maxWin = d->prm.cornerRefinementWinSize;
        sp.maxIter = d->prm.cornerRefinementMaxIterations; sp.relWin = d->prm.relativeCornerRefinmentWinSize; sp.eps = d->prm.cornerRefinementMinAccuracy;
        float *c2 = d->d_corners2 + (size_t)b0 * K * 8;
        const float scale_init = (float)pl.W[s.closestIdx] / (float)pl.segW;
        if (s.closestIdx == 0) {
            sp.W = pl.W[0]; sp.H = pl.H[0]; sp.pitch = pl.pitch[0]; sp.frame_stride = pl.frame_stride[0];
            sp.fixedWin = 3; sp.refine = 0; sp.mul0 = scale_init; sp.mul1 = 1.f;
            launch_subpix(d, st, pl.base[0], fa.fo0.n_accepted, fa.fo0.corners, c2, nb, sp);
        } else {
            for (int idx = s.closestIdx - 1; idx >= 0; --idx) {
                const bool first = idx == s.closestIdx - 1;
                sp.W = pl.W[idx]; sp.H = pl.H[idx]; sp.pitch = pl.pitch[idx]; sp.frame_stride = pl.frame_stride[idx];
                sp.fixedWin = std::max(pl.W[idx], pl.H[idx]) > 1080 ? 5 : 3; sp.refine = 1;
                sp.mul0 = first ? scale_init : 2.f; sp.mul1 = first ? 2.f : 1.f;
                launch_subpix(d, st, pl.base[idx], fa.fo0.n_accepted, first ? fa.fo0.corners : c2, c2, nb, sp);
            }
        }
        corners = c2;
    } else if (d->prm.cornerRefinementMethod == 1) {
        SubpixParams sp;
        sp.W = g.W; sp.H = g.H; sp.pitch = s.pitch; sp.frame_stride = s.frame_stride; sp.max_markers = d->max_markers;
        sp.markerSize = d->dict.markerSize; sp.borderBits = d->prm.markerBorderBits; sp.maxWin = d->prm.cornerRefinementWinSize;
        sp.maxIter = d->prm.cornerRefinementMaxIterations; sp.relWin = d->prm.relativeCornerRefinmentWinSize; sp.eps = d->prm.cornerRefinementMinAccuracy;
        sp.fixedWin = 0; sp.refine = 1; sp.mul0 = 1.f; sp.mul1 = 1.f;
        float *c2 = d->d_corners2 + (size_t)b0 * K * 8;
        launch_subpix(d, st, s.gray, fa.fo0.n_accepted, fa.fo0.corners, c2, nb, sp);
        corners = c2;
    } else if (d->prm.cornerRefinementMethod == 2) {
        float *c2 = d->d_corners2 + (size_t)b0 * K * 8;
        k_refine_contour<<<(nb * (int)K * 32 + 127) / 128, 128, 0, st>>>(fa.fo0.corners, fa.fo0.n_accepted, c2, d->d_surv_count + fs0, d->d_pts_off + fs0 * g.surv_cap,
                                                                        d->d_pts + fs0 * (size_t)g.pts_cap, d->d_quad_ok + fs0 * g.surv_cap, d->d_quad_xy + fs0 * g.surv_cap * 8,
                                                                        d->d_quad_len + fs0 * g.surv_cap, g, nb, d->max_markers);
        d->launches++;
        corners = c2;
    }
    stage_mark(d, s, ST_POSE);
    long long *pose_marks = nullptr;
#ifdef B2A_DEBUG_TAPS
    static long long *pose_marks_buf = nullptr;
    if (std::getenv("B2A_POSE_MARKS") && s.sb == 0) {
        if (!pose_marks_buf) cudaMalloc(&pose_marks_buf, 32 * sizeof(long long));
        cudaMemsetAsync(pose_marks_buf, 0, 32 * sizeof(long long), st);
        pose_marks = pose_marks_buf;
    }
#endif
    if (cam) {
        k_pose<<<(nb * (int)K * 32 + POSE_THREADS - 1) / POSE_THREADS, POSE_THREADS, 0, st>>>(corners, fa.fo0.n_accepted, nb, d->max_markers, to_camera(cam), cam->marker_length,
                                                        d->d_rvecs + (size_t)b0 * K * 3, d->d_tvecs + (size_t)b0 * K * 3, pose_marks);
        d->launches++;
#ifdef B2A_DEBUG_TAPS
        if (pose_marks) {
            long long hm[32];
            cudaStreamSynchronize(st);
            cudaMemcpy(hm, pose_marks, sizeof(hm), cudaMemcpyDeviceToHost);
            std::fprintf(stderr, "k_pose marker 0 cycles (init, first rows, then per LM iteration):");
            for (int i = 1; i < 32 && hm[i]; ++i) std::fprintf(stderr, " %lld", hm[i] - hm[i - 1]);
            std::fprintf(stderr, "\n");
        }
#endif
    }
    stage_mark(d, s, ST_D2H);
    {
        const size_t o = (size_t)b0 * K;
        ExportPtrs ep;
        ep.n_acc = fa.fo0.n_accepted; ep.n_rej = fa.fo0.n_rejected; ep.status = d->d_status + b0; ep.ids = fa.fo0.ids;
        ep.corners = corners; ep.rejected = fa.fo0.rejected; ep.rvecs = d->d_rvecs + o * 3; ep.tvecs = d->d_tvecs + o * 3;
        ep.h_nacc = d->h_nacc + b0; ep.h_nrej = d->h_nrej + b0; ep.h_status = d->h_status + b0; ep.h_ids = d->h_ids + o;
        ep.h_corners = d->h_corners + o * 8; ep.h_rejected = d->h_rejected + o * 8; ep.h_rvecs = d->h_rvecs + o * 3; ep.h_tvecs = d->h_tvecs + o * 3;
        k_export<<<nb, 128, 0, st>>>(ep, d->max_markers, cam ? 1 : 0);
        d->launches++;
    }
    if (s.timed) cudaEventRecord(d->ev[ST_COUNT], st);
    return launch_err("back-end kernels");
}

// mode: 0 full pipeline, 1 stop after the front end (taps), 2 stop after grouping (candidate tap)
// post: work the caller appends on the handle's stream after the sub-batches have joined
// enqueue_pipeline returns as soon as everything is queued; finish_pipeline is the one synchronisation of a call
constexpr int GRAPH_MAX_BATCH = 4, GRAPH_CACHE = 8;
static int enqueue_body(b2a_detector *d, const b2a_frames *f, const b2a_camera *cam, int mode, int walk_max_len, std::vector<Sub> *subs_out);

// the call as one graph launch: frames into d_in (plain copies on the handle's stream), then the captured kernels of this
// (batch, shape, camera); the first call of a kind captures and instantiates.  A capture that fails switches the handle back
// to plain launches for good.
static int enqueue_graph(b2a_detector *d, const b2a_frames *f, const b2a_camera *cam)
{
    const int B = f->batch, W = f->width, H = f->height;
    cudaStream_t s0 = d->stream;
    const size_t rowbytes = (size_t)W * f->channels, dpitch = f->channels == 1 ? ((rowbytes + 3) & ~(size_t)3) : rowbytes;
    const size_t in_pitch = f->row_stride ? f->row_stride : rowbytes, in_frame = f->frame_stride ? f->frame_stride : in_pitch * H;
    const cudaMemcpyKind kind = f->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (in_frame == in_pitch * H && in_pitch == rowbytes && dpitch == rowbytes) CU(cudaMemcpyAsync(d->d_in, f->data, rowbytes * H * B, kind, s0));
    else if (in_frame == in_pitch * H) CU(cudaMemcpy2DAsync(d->d_in, dpitch, f->data, in_pitch, rowbytes, (size_t)H * B, kind, s0));
    else for (int b = 0; b < B; ++b) CU(cudaMemcpy2DAsync(d->d_in + (size_t)b * dpitch * H, dpitch, f->data + (size_t)b * in_frame, in_pitch, rowbytes, H, kind, s0));
    b2a_frames fg = *f;
    fg.data = d->d_in; fg.on_device = 1; fg.row_stride = dpitch; fg.frame_stride = dpitch * H;
    b2a_camera cz;
    std::memset(&cz, 0, sizeof(cz));
    if (cam) { std::memcpy(cz.K, cam->K, sizeof(cz.K)); std::memcpy(cz.D, cam->D, sizeof(cz.D)); cz.nD = cam->nD; cz.marker_length = cam->marker_length; }
    b2a_detector::GraphEntry *e = nullptr;
    for (auto &g : d->graphs)
        if (g.B == B && g.W == W && g.H == H && g.channels == f->channels && g.n_streams == d->n_streams && g.has_cam == (cam != nullptr) &&
            g.pipelined == d->pipelined && std::memcmp(&g.cam, &cz, sizeof(cz)) == 0) { e = &g; break; }
    if (!e) {
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal);
        int rc = B2A_ERR_CUDA;
        if (ce == cudaSuccess) {
            d->capturing = true;
            d->launches = 0;
            rc = enqueue_body(d, &fg, cam, 0, 0, nullptr);
            d->capturing = false;
            ce = cudaStreamEndCapture(s0, &graph);
        }
        cudaGraphExec_t exec = nullptr;
        if (rc == B2A_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != B2A_OK || ce != cudaSuccess || !exec) {
            cudaGetLastError();
            d->graph_failed = true;                          // plain launches from now on (the frames are already in d_in)
            d->launches = 0;
            return enqueue_body(d, &fg, cam, 0, 0, nullptr);
        }
        if ((int)d->graphs.size() >= GRAPH_CACHE) {
            size_t lru = 0;
            for (size_t i = 1; i < d->graphs.size(); ++i) if (d->graphs[i].last_use < d->graphs[lru].last_use) lru = i;
            cudaGraphExecDestroy(d->graphs[lru].exec);
            d->graphs.erase(d->graphs.begin() + (long)lru);
        }
        d->graphs.push_back(b2a_detector::GraphEntry{B, W, H, f->channels, d->n_streams, cam != nullptr, d->pipelined, cz, exec, d->launches, 0});
        e = &d->graphs.back();
    }
    e->last_use = ++d->graph_tick;
    CU(cudaGraphLaunch(e->exec, s0));
    d->launches = e->launches;
    return B2A_OK;
}

static int enqueue_pipeline(b2a_detector *d, const b2a_frames *f, const b2a_camera *cam, int mode, int walk_max_len, std::vector<Sub> *subs_out,
                            const std::function<int(cudaStream_t)> *post = nullptr)
{
    TRY(check_frames(d, f));
    CU(cudaSetDevice(d->device));
    const int B = f->batch, W = f->width, H = f->height;
    std::memset(d->ev_used, 0, sizeof(d->ev_used));
    d->launches = 0;
    d->timed_mode = mode;
    d->last_call_batch = mode == 0 ? B : 0; d->last_call_pose = cam != nullptr;
    cudaStream_t s0 = d->stream;
    if (W != d->lastW || H != d->lastH || B > d->lastB) {       // padding words / rows of the masks must be zero
        CU(cudaMemsetAsync(d->d_masks, 0, d->masks_words * sizeof(uint32_t), s0));
        d->lastW = W; d->lastH = H; d->lastB = std::max(B, d->lastB);
    }
    static const bool graph_env = !(std::getenv("B2A_GRAPH") && std::atoi(std::getenv("B2A_GRAPH")) == 0);
    if (graph_env && d->use_graph && !d->graph_failed && mode == 0 && !subs_out && B <= GRAPH_MAX_BATCH && !sync_debug()) {
        int rc = enqueue_graph(d, f, cam);
        if (rc != B2A_OK) return rc;
    } else TRY(enqueue_body(d, f, cam, mode, walk_max_len, subs_out));
    if (post) TRY((*post)(s0));
    return B2A_OK;
}

static int enqueue_body(b2a_detector *d, const b2a_frames *f, const b2a_camera *cam, int mode, int walk_max_len, std::vector<Sub> *subs_out)
{
    const int B = f->batch;
    cudaStream_t s0 = d->stream;
    const size_t FSmax = (size_t)d->cfg.max_batch * d->nScales;
    CU(cudaMemsetAsync(d->d_counters, 0, (d->n_sub_max + 3 * FSmax + d->cfg.max_batch) * sizeof(int), s0));
    CU(cudaMemsetAsync(d->d_counters2, 0, d->n_sub_max * sizeof(unsigned), s0));
    // sub-batches
    // frames in HBM: two sub-batches overlap the latency-bound back end of one with the front end of the other (measured best:
    // 29.5 k frames/s against 26.4 k with one and 28.4 k with four); host frames: four, so that kernels run under the PCIe copy
    // submitted (pipelined) batches overlap their PCIe copy with the OTHER context's kernels, so two even sub-batches are enough there
    // and the kernels run at their better batch size (measured on C2: 25.3 k frames/s with 2, 22.2 k with 4, 19.0 k with 8)
    const int want_streams = d->n_streams > 0 ? d->n_streams : ((f->on_device || d->pipelined) ? 2 : 4);
    int nsub = std::max(1, std::min(std::min(want_streams, d->n_sub_max), B));
    // sub-batch boundaries.  Frames that still have to cross PCIe are cut unevenly: a small first sub-batch lets the
    // kernels start early and a small last one shortens the tail that nothing overlaps (B2A_SPLIT="2,6,8,8,6,2" overrides)
    std::vector<int> bounds;
    static const char *split_env = std::getenv("B2A_SPLIT");
    if (split_env) {
        std::vector<int> w;
        for (const char *p = split_env; *p;) { w.push_back(std::max(1, std::atoi(p))); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
        if ((int)w.size() > d->n_sub_max) w.resize(d->n_sub_max);
        int tot = 0; for (int v : w) tot += v;
        bounds.push_back(0);
        int acc = 0;
        for (size_t i = 0; i < w.size(); ++i) { acc += w[i]; const int b = (int)((long long)B * acc / tot); if (b > bounds.back()) bounds.push_back(b); }
        if (bounds.back() != B) bounds.push_back(B);
    } else if (!f->on_device && nsub >= 4 && B >= 4 * nsub) {
        // weights 1 : 2 : ... : 2 : 1
        const int tot = 2 * nsub - 2;
        bounds.push_back(0);
        for (int i = 1; i <= nsub; ++i) { const int acc = (i == nsub) ? tot : 2 * i - 1; bounds.push_back((int)((long long)B * acc / tot)); }
    } else {
        for (int i = 0; i <= nsub; ++i) bounds.push_back((int)((long long)B * i / nsub));
    }
    nsub = (int)bounds.size() - 1;
    std::vector<Sub> subs(nsub);
    for (int i = 0; i < nsub; ++i) {
        Sub &s = subs[i];
        s.sb = i; s.b0 = bounds[i]; s.nb = bounds[i + 1] - bounds[i];
        s.st = d->streams[i]; s.timed = (i == 0) && !d->capturing;       // stage events cannot be read back from a replayed graph
    }
    CU(cudaEventRecord(d->ev_fork, s0));
    for (int i = 1; i < nsub; ++i) CU(cudaStreamWaitEvent(subs[i].st, d->ev_fork, 0));
#ifdef B2A_DEBUG_TAPS
    // B2A_TIMELINE=1: per sub-batch, when its frames were in HBM and when its results were out (ms after the call began)
    static const bool timeline = std::getenv("B2A_TIMELINE") != nullptr;
    static cudaEvent_t tl_ev[1 + 2 * b2a_detector::MAX_SUB] = {};
    if (timeline) {
        if (!tl_ev[0]) for (auto &e : tl_ev) cudaEventCreate(&e);
        cudaEventRecord(tl_ev[0], s0);
    }
#endif
    for (int i = 0; i < nsub; ++i) {
#ifdef B2A_DEBUG_TAPS
        subs[i].tl_after_h2d = timeline ? tl_ev[1 + 2 * i] : nullptr;
#endif
        TRY(run_front(d, f, subs[i], walk_max_len));
        if (mode != 1) TRY(run_back(d, subs[i], cam, mode == 2));
#ifdef B2A_DEBUG_TAPS
        if (timeline) cudaEventRecord(tl_ev[2 + 2 * i], subs[i].st);
#endif
    }
    for (int i = 1; i < nsub; ++i) { CU(cudaEventRecord(d->ev_join[i], subs[i].st)); CU(cudaStreamWaitEvent(s0, d->ev_join[i], 0)); }
#ifdef B2A_DEBUG_TAPS
    if (timeline && mode == 0) {
        cudaStreamSynchronize(s0);
        std::fprintf(stderr, "timeline (ms): ");
        for (int i = 0; i < nsub; ++i) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, tl_ev[0], tl_ev[1 + 2 * i]); cudaEventElapsedTime(&b, tl_ev[0], tl_ev[2 + 2 * i]);
            std::fprintf(stderr, "[%d frames: in %.3f out %.3f] ", subs[i].nb, a, b);
        }
        std::fprintf(stderr, "\n");
    }
#endif
    if (subs_out) *subs_out = subs;
    return B2A_OK;
}

static int finish_pipeline(b2a_detector *d)
{
    CU(cudaSetDevice(d->device));
    CU(cudaStreamSynchronize(d->stream));
    if (d->timed_mode == 0) {
        int prev = -1;
        for (int i = 0; i < ST_COUNT; ++i) d->stage_ms[i] = 0.f;
        for (int i = 0; i <= ST_COUNT; ++i) {
            if (i < ST_COUNT && !d->ev_used[i]) continue;
            if (prev >= 0) cudaEventElapsedTime(&d->stage_ms[prev], d->ev[prev], d->ev[i]);
            prev = i;
        }
    }
    return B2A_OK;
}

static int run_pipeline(b2a_detector *d, const b2a_frames *f, const b2a_camera *cam, int mode, int walk_max_len, std::vector<Sub> *subs_out,
                        const std::function<int(cudaStream_t)> *post = nullptr)
{
    if (d && d->in_flight) return set_err(B2A_ERR_INVALID, "a submitted batch is still in flight on this handle (b2a_detect_pose_wait first)");
#ifdef B2A_DEBUG_TAPS
    static const bool host_time = std::getenv("B2A_HOSTTIME") != nullptr;     // debug: host time spent enqueueing against the whole call
    const auto t_call = std::chrono::steady_clock::now();
#endif
    TRY(enqueue_pipeline(d, f, cam, mode, walk_max_len, subs_out, post));
#ifdef B2A_DEBUG_TAPS
    const auto t_enq = std::chrono::steady_clock::now();
#endif
    TRY(finish_pipeline(d));
#ifdef B2A_DEBUG_TAPS
    if (host_time) {
        static double acc_e = 0, acc_t = 0; static int cnt = 0;
        const auto t_end = std::chrono::steady_clock::now();
        acc_e += std::chrono::duration<double, std::milli>(t_enq - t_call).count(); acc_t += std::chrono::duration<double, std::milli>(t_end - t_call).count();
        if (++cnt % 20 == 0) { std::fprintf(stderr, "host: enqueue %.3f ms of %.3f ms per call (%d launches)\n", acc_e / 20, acc_t / 20, d->launches); acc_e = acc_t = 0; }
    }
#endif
    return B2A_OK;
}

static int fill_out(b2a_detector *d, int B, bool pose, b2a_detections *out)
{
    out->batch = B; out->max_markers = d->max_markers;
    out->n_accepted = d->h_nacc; out->n_rejected = d->h_nrej; out->corners = d->h_corners; out->ids = d->h_ids;
    out->rejected = d->h_rejected; out->rvecs = pose ? d->h_rvecs : nullptr; out->tvecs = pose ? d->h_tvecs : nullptr;
    out->status = d->h_status;
    for (int b = 0; b < B; ++b) if (d->h_status[b] != 0) return set_err(B2A_ERR_CAPACITY, "an internal list overflowed (see b2a_detections.status)");
    return B2A_OK;
}

// compact records of a call's detections: what travels when results are gathered from several GPUs / processes
extern "C" int b2a_pack_detections(const b2a_detections *det, void *out, size_t cap, size_t *n_bytes)
{
    if (!det || !n_bytes || (cap > 0 && !out)) return set_err(B2A_ERR_INVALID, "null argument");
    const int B = det->batch, K = det->max_markers;
    const bool pose = det->rvecs && det->tvecs;
    size_t need = 16;
    for (int b = 0; b < B; ++b) need += 8 + (size_t)det->n_accepted[b] * (4 + 32 + (pose ? 48 : 0)) + (size_t)det->n_rejected[b] * 32;
    *n_bytes = need;
    if (need > cap) return set_err(B2A_ERR_CAPACITY, "packed detections do not fit the buffer");
    uint8_t *p = (uint8_t *)out;
    const int32_t hdr[4] = {0x42324144 /* "DA2B" */, B, pose ? 1 : 0, K};
    std::memcpy(p, hdr, 16); p += 16;
    for (int b = 0; b < B; ++b) {
        const int na = det->n_accepted[b], nr = det->n_rejected[b];
        const int32_t cnt[2] = {na, nr};
        std::memcpy(p, cnt, 8); p += 8;
        std::memcpy(p, det->ids + (size_t)b * K, (size_t)na * 4); p += (size_t)na * 4;
        std::memcpy(p, det->corners + (size_t)b * K * 8, (size_t)na * 32); p += (size_t)na * 32;
        if (pose) {
            std::memcpy(p, det->rvecs + (size_t)b * K * 3, (size_t)na * 24); p += (size_t)na * 24;
            std::memcpy(p, det->tvecs + (size_t)b * K * 3, (size_t)na * 24); p += (size_t)na * 24;
        }
        std::memcpy(p, det->rejected + (size_t)b * K * 8, (size_t)nr * 32); p += (size_t)nr * 32;
    }
    return B2A_OK;
}

extern "C" int b2a_detector_last_detections(b2a_detector *d, b2a_detections *out)
{
    if (!d || !out) return set_err(B2A_ERR_INVALID, "null argument");
    if (d->in_flight) return set_err(B2A_ERR_INVALID, "a submitted batch is still in flight on this handle");
    if (d->lastB <= 0 || d->last_call_batch <= 0) return set_err(B2A_ERR_INVALID, "no completed call on this handle yet");
    return fill_out(d, d->last_call_batch, d->last_call_pose, out);
}

extern "C" int b2a_detect(b2a_detector *d, const b2a_frames *frames, b2a_detections *out)
{
    if (!out) return set_err(B2A_ERR_INVALID, "null output");
    TRY(run_pipeline(d, frames, nullptr, 0, 0, nullptr));
    return fill_out(d, frames->batch, false, out);
}

extern "C" int b2a_detect_pose(b2a_detector *d, const b2a_frames *frames, const b2a_camera *cam, b2a_detections *out)
{
    if (!out || !cam) return set_err(B2A_ERR_INVALID, "null argument");
    if (!(cam->marker_length > 0)) return set_err(B2A_ERR_INVALID, "markerLength <= 0");
    TRY(run_pipeline(d, frames, cam, 0, 0, nullptr));
    return fill_out(d, frames->batch, true, out);
}

// post(context, stream): work appended behind the batch on the context that runs it
static int submit_impl(b2a_detector *d, const b2a_frames *frames, const b2a_camera *cam, int *ticket,
                       const std::function<int(b2a_detector *, cudaStream_t)> *post)
{
    if (!d || !frames || !ticket) return set_err(B2A_ERR_INVALID, "null argument");
    if (cam && !(cam->marker_length > 0)) return set_err(B2A_ERR_INVALID, "markerLength <= 0");
    b2a_detector *t = d;
    const int slot = (int)(d->next_ticket % (unsigned)d->n_ctx);
    if (slot > 0) {
        if (!d->more[slot - 1]) {                              // further contexts are created on first use
            TRY(b2a_detector_create(&d->cfg, &d->dict, &d->prm, &d->more[slot - 1]));
        }
        t = d->more[slot - 1];
        t->n_streams = d->n_streams; t->use_graph = d->use_graph;
    }
    if (t->in_flight) return set_err(B2A_ERR_INVALID, "all contexts of this handle are in flight (b2a_detect_pose_wait first)");
    t->pipelined = true;
    std::function<int(cudaStream_t)> bound;
    if (post) bound = [&](cudaStream_t st) { return (*post)(t, st); };
    const int rc_enq = enqueue_pipeline(t, frames, cam, 0, 0, nullptr, post ? &bound : nullptr);
    t->pipelined = false;
    TRY(rc_enq);
    t->in_flight = true; t->pending_pose = cam != nullptr; t->pending_batch = frames->batch;
    *ticket = (int)d->next_ticket++;
    return B2A_OK;
}

extern "C" int b2a_detector_set_graph(b2a_detector *d, int on)
{
    if (!d) return set_err(B2A_ERR_INVALID, "null argument");
    d->use_graph = on != 0;
    return B2A_OK;
}

extern "C" int b2a_detect_pose_submit(b2a_detector *d, const b2a_frames *frames, const b2a_camera *cam, int *ticket)
{
    return submit_impl(d, frames, cam, ticket, nullptr);
}

extern "C" int b2a_detect_pose_wait(b2a_detector *d, int ticket, b2a_detections *out)
{
    if (!d || !out) return set_err(B2A_ERR_INVALID, "null argument");
    const int slot = ticket < 0 ? 0 : ticket % d->n_ctx;
    b2a_detector *t = slot ? d->more[slot - 1] : d;
    if (ticket < 0 || !t || !t->in_flight || (unsigned)ticket + (unsigned)d->n_ctx < d->next_ticket || (unsigned)ticket >= d->next_ticket)
        return set_err(B2A_ERR_INVALID, "no batch in flight under this ticket");
    t->in_flight = false;
    TRY(finish_pipeline(t));
    if (t != d) { d->launches = t->launches; std::memcpy(d->stage_ms, t->stage_ms, sizeof(d->stage_ms)); }
    return fill_out(t, t->pending_batch, t->pending_pose, out);
}

// ------------------------------------------------------------------------------------------------
// overlay: cv::aruco::drawDetectedMarkers (aruco_slam.cpp:319).  One CTA per image; the markers are drawn one after the other (a later
// marker's lines may cross an earlier label), inside a marker the four lines are walked by four threads, the corner stamp is a pixel
// per thread and the label a pixel per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_draw_markers(OverlayImage im, OverlayTables t, const float *__restrict__ corners, const int32_t *__restrict__ ids, const int32_t *__restrict__ n_dev, int n,
               uchar3 border)
{
    if (n_dev) n = *n_dev;
    const uint8_t bc[3] = {border.x, border.y, border.z};
    uint8_t text[3], corner[3];
    overlay_colours(bc, text, corner);
    for (int i = 0; i < n; ++i) {
        const float *q = corners + 8 * i;
        int px[4], py[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { px[j] = overlay_round(q[2 * j]); py[j] = overlay_round(q[2 * j + 1]); }
        if (threadIdx.x < 4) { const int j = threadIdx.x; overlay_line(im, px[j], py[j], px[(j + 1) & 3], py[(j + 1) & 3], bc); }
        __syncthreads();
        if (threadIdx.x < 121) overlay_stamp_pixel(im, t, px[0], py[0], (int)threadIdx.x % 11 - 5, (int)threadIdx.x / 11 - 5, corner);
        __syncthreads();
        if (ids) {
            int ox, oy;
            overlay_text_origin(q, ox, oy);
            const int id = ids[i];
            for (int k = threadIdx.x; k < 16 * 72; k += blockDim.x) {
                const int tx = k % 72, ty = k / 72 - 13;
                if (overlay_text_bit(t, id, tx, ty)) overlay_put(im, ox + tx, oy + ty, text);
            }
            __syncthreads();
        }
    }
}

extern "C" int b2a_draw_detected_markers(b2a_detector *d, uint8_t *image, int width, int height, int channels, size_t row_stride,
                                         const float *corners, const int32_t *ids, int n, const uint8_t border_bgr[3])
{
    if (!d || !image || (n > 0 && !corners)) return set_err(B2A_ERR_INVALID, "null argument");
    if (channels != 1 && channels != 3) return set_err(B2A_ERR_INVALID, "channels must be 1 or 3");            // CV_Assert in drawDetectedMarkers
    if (width <= 0 || height <= 0 || (size_t)width * height > (size_t)d->cfg.max_width * d->cfg.max_height) return set_err(B2A_ERR_INVALID, "image size outside the handle's maximum");
    if (n < 0 || n > d->max_markers) return set_err(B2A_ERR_INVALID, "more markers than max_markers");
    if (d->in_flight) return set_err(B2A_ERR_INVALID, "a submitted batch is still in flight on this handle");
    if (ids) for (int i = 0; i < n; ++i) if (ids[i] < 0 || ids[i] > 9999) return set_err(B2A_ERR_UNSUPPORTED, "marker id outside 0 .. 9999");
    if (n == 0) return B2A_OK;
    CU(cudaSetDevice(d->device));
    static const OverlayTables tables = make_overlay_tables();
    const size_t rowbytes = (size_t)width * channels, pitch = row_stride ? row_stride : rowbytes;
    cudaStream_t st = d->stream;
    // the image visits the frame buffer of the handle; corners / ids use the output arrays of the first frame slot
    CU(cudaMemcpy2DAsync(d->d_in, rowbytes, image, pitch, rowbytes, height, cudaMemcpyHostToDevice, st));
    float *dc = d->d_corners2;
    int32_t *di = (int32_t *)d->d_wM;
    CU(cudaMemcpyAsync(dc, corners, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    if (ids) CU(cudaMemcpyAsync(di, ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    OverlayImage im{d->d_in, width, height, channels, rowbytes};
    const uint8_t dflt[3] = {0, 255, 0};
    const uint8_t *b = border_bgr ? border_bgr : dflt;
    k_draw_markers<<<1, 256, 0, st>>>(im, tables, dc, ids ? di : nullptr, nullptr, n, make_uchar3(b[0], b[1], b[2]));
    CU(cudaMemcpy2DAsync(image, pitch, d->d_in, rowbytes, rowbytes, height, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return launch_err("k_draw_markers");
}

// ------------------------------------------------------------------------------------------------
// several GPUs of one box from one process: frames are independent, so a batch is cut into contiguous blocks, one per
// device, each block goes through its device's handle from its own host thread, and only the detections are gathered -- on the
// host, in frame order.  No collective, no peer traffic.
// ------------------------------------------------------------------------------------------------
struct b2a_multi {
    std::vector<b2a_detector *> dets;
    int max_batch = 0, max_markers = 0;
    std::vector<int32_t> nacc, nrej, status, ids;
    std::vector<float> corners, rejected;
    std::vector<double> rvecs, tvecs;
};

extern "C" void b2a_multi_destroy(b2a_multi *m)
{
    if (!m) return;
    for (b2a_detector *d : m->dets) b2a_detector_destroy(d);
    delete m;
}

extern "C" int b2a_multi_create(const int *devices, int n_devices, const b2a_detector_config *cfg, const b2a_dictionary *dict,
                                const b2a_detector_params *params, b2a_multi **out)
{
    if (!devices || n_devices <= 0 || !cfg || !dict || !out) return set_err(B2A_ERR_INVALID, "null argument");
    b2a_multi *m = new b2a_multi();
    m->max_batch = cfg->max_batch;
    const int per = (cfg->max_batch + n_devices - 1) / n_devices;
    for (int g = 0; g < n_devices; ++g) {
        b2a_detector_config c = *cfg;
        c.device = devices[g]; c.max_batch = std::max(per, 1);
        b2a_detector *d = nullptr;
        const int rc = b2a_detector_create(&c, dict, params, &d);
        if (rc != B2A_OK) { const std::string keep = g_err; b2a_multi_destroy(m); g_err = keep; return rc; }
        m->dets.push_back(d);
    }
    m->max_markers = m->dets[0]->max_markers;
    const size_t BK = (size_t)cfg->max_batch * m->max_markers;
    m->nacc.resize(cfg->max_batch); m->nrej.resize(cfg->max_batch); m->status.resize(cfg->max_batch); m->ids.resize(BK);
    m->corners.resize(BK * 8); m->rejected.resize(BK * 8); m->rvecs.resize(BK * 3); m->tvecs.resize(BK * 3);
    *out = m;
    return B2A_OK;
}

extern "C" int b2a_multi_num_devices(const b2a_multi *m) { return m ? (int)m->dets.size() : 0; }

extern "C" int b2a_multi_detect_pose(b2a_multi *m, const b2a_frames *f, const b2a_camera *cam, b2a_detections *out)
{
    if (!m || !f || !out || !f->data) return set_err(B2A_ERR_INVALID, "null argument");
    if (f->on_device) return set_err(B2A_ERR_INVALID, "the frames of a multi-device call live in host memory");
    if (f->batch <= 0 || f->batch > m->max_batch) return set_err(B2A_ERR_INVALID, "batch outside [1, max_batch]");
    if (cam && !(cam->marker_length > 0)) return set_err(B2A_ERR_INVALID, "markerLength <= 0");
    const int G = (int)m->dets.size(), B = f->batch, K = m->max_markers;
    const size_t in_pitch = f->row_stride ? f->row_stride : (size_t)f->width * f->channels;
    const size_t in_frame = f->frame_stride ? f->frame_stride : in_pitch * f->height;
    std::vector<int> rcs(G, B2A_OK);
    std::vector<std::string> errs(G);
    std::vector<b2a_detections> dets(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        const int b0 = (int)((long long)B * g / G), b1 = (int)((long long)B * (g + 1) / G);
        if (b1 <= b0) continue;
        th.emplace_back([&, g, b0, b1]() {
            b2a_frames fr = *f;
            fr.data = f->data + (size_t)b0 * in_frame; fr.batch = b1 - b0; fr.row_stride = in_pitch; fr.frame_stride = in_frame;
            rcs[g] = cam ? b2a_detect_pose(m->dets[g], &fr, cam, &dets[g]) : b2a_detect(m->dets[g], &fr, &dets[g]);
            if (rcs[g] != B2A_OK) errs[g] = g_err;             // the error text is thread local
        });
    }
    for (std::thread &t : th) t.join();
    // gather in frame order (only the filled entries move)
    int rc = B2A_OK;
    for (int g = 0; g < G; ++g) {
        const int b0 = (int)((long long)B * g / G), b1 = (int)((long long)B * (g + 1) / G);
        if (b1 <= b0) continue;
        if (rcs[g] != B2A_OK && rcs[g] != B2A_ERR_CAPACITY) return set_err(rcs[g], "device block " + std::to_string(g) + ": " + errs[g]);
        if (rcs[g] == B2A_ERR_CAPACITY) { rc = B2A_ERR_CAPACITY; g_err = errs[g]; }
        const b2a_detector *d = m->dets[g];
        for (int b = b0; b < b1; ++b) {
            const int l = b - b0, na = d->h_nacc[l], nr = d->h_nrej[l];
            m->nacc[b] = na; m->nrej[b] = nr; m->status[b] = d->h_status[l];
            std::memcpy(&m->ids[(size_t)b * K], d->h_ids + (size_t)l * K, (size_t)na * sizeof(int32_t));
            std::memcpy(&m->corners[(size_t)b * K * 8], d->h_corners + (size_t)l * K * 8, (size_t)na * 8 * sizeof(float));
            std::memcpy(&m->rejected[(size_t)b * K * 8], d->h_rejected + (size_t)l * K * 8, (size_t)nr * 8 * sizeof(float));
            if (cam) {
                std::memcpy(&m->rvecs[(size_t)b * K * 3], d->h_rvecs + (size_t)l * K * 3, (size_t)na * 3 * sizeof(double));
                std::memcpy(&m->tvecs[(size_t)b * K * 3], d->h_tvecs + (size_t)l * K * 3, (size_t)na * 3 * sizeof(double));
            }
        }
    }
    out->batch = B; out->max_markers = K;
    out->n_accepted = m->nacc.data(); out->n_rejected = m->nrej.data(); out->status = m->status.data(); out->ids = m->ids.data();
    out->corners = m->corners.data(); out->rejected = m->rejected.data();
    out->rvecs = cam ? m->rvecs.data() : nullptr; out->tvecs = cam ? m->tvecs.data() : nullptr;
    return rc;
}

extern "C" int b2a_detector_set_inflight(b2a_detector *d, int n)
{
    if (!d || n < 1 || n > b2a_detector::MAX_CTX) return set_err(B2A_ERR_INVALID, "in-flight batches must be 1 .. 8");
    if (d->in_flight) return set_err(B2A_ERR_INVALID, "a submitted batch is still in flight on this handle");
    for (b2a_detector *m : d->more) if (m && m->in_flight) return set_err(B2A_ERR_INVALID, "a submitted batch is still in flight on this handle");
    d->n_ctx = n;
    d->next_ticket = 0;
    return B2A_OK;
}

extern "C" int b2a_detector_set_streams(b2a_detector *d, int n)
{
    if (!d || n < 0) return set_err(B2A_ERR_INVALID, "streams must be >= 0 (0 = automatic)");
    d->n_streams = std::min(n, d->n_sub_max);
    return B2A_OK;
}

// ---- refineDetectedMarkers ----------------------------------------------------------------------------------------------
extern "C" void b2a_default_refine_params(b2a_refine_params *p) { p->minRepDistance = 10.f; p->errorCorrectionRate = 3.f; p->checkAllOrders = 1; }

// the inner bits of every (rejected candidate, corner order) of one frame: the identification kernels run on a work list
// written by the host into frame slot 0 (the detector's results of its last call are overwritten)
static int refine_extract_codes(b2a_detector *d, const b2a_frames *f, const float *rejected, int n_rej, std::vector<unsigned long long> &codes)
{
    const int W = f->width, H = f->height, nw = 4 * n_rej;
    if (nw > d->max_cand) return set_err(B2A_ERR_CAPACITY, "4 x rejected candidates exceed max_candidates");
    cudaStream_t st = d->stream;
    const size_t in_pitch = f->row_stride ? f->row_stride : (size_t)W * f->channels;
    const uint8_t *src = f->data;
    size_t src_pitch = in_pitch;
    if (!f->on_device) {
        const size_t rowbytes = (size_t)W * f->channels;
        CU(cudaMemcpy2DAsync(d->d_in, rowbytes, f->data, in_pitch, rowbytes, H, cudaMemcpyHostToDevice, st));
        src = d->d_in; src_pitch = rowbytes;
    }
    const uint8_t *gray = src;
    size_t gpitch = src_pitch;
    if (f->channels == 3) {
        k_bgr2gray<<<d->num_sms * 4, 256, 0, st>>>(src, src_pitch, src_pitch * H, d->d_gray, d->gray_pitch, d->gray_pitch * H, W, H, 1);
        gray = d->d_gray; gpitch = d->gray_pitch;
    }
    std::vector<float> wq((size_t)nw * 8);
    for (int j = 0; j < n_rej; ++j)
        for (int c = 0; c < 4; ++c)
            for (int k = 0; k < 4; ++k) { wq[((size_t)j * 4 + c) * 8 + 2 * k] = rejected[j * 8 + 2 * ((c + k) & 3)]; wq[((size_t)j * 4 + c) * 8 + 2 * k + 1] = rejected[j * 8 + 2 * ((c + k) & 3) + 1]; }
    int counters[8] = {0, 0, nw, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(d->fs0.wq, wq.data(), wq.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d->fs0.counters, counters, sizeof(counters), cudaMemcpyHostToDevice, st));
    FrameArrays fa;
    fa.fs0 = d->fs0; fa.fo0 = d->fo0; fa.fo0.status = d->d_status;
    fa.surv_count = d->d_surv_count; fa.quad_ok = d->d_quad_ok; fa.quad_xy = d->d_quad_xy; fa.quad_len = d->d_quad_len; fa.marks = nullptr;
    IdentParams ip;
    ip.markerSize = d->dict.markerSize; ip.borderBits = d->prm.markerBorderBits; ip.cellSize = d->prm.perspectiveRemovePixelPerCell;
    ip.cellMargin = (int)(d->prm.perspectiveRemoveIgnoredMarginPerCell * ip.cellSize);
    ip.nMarkers = d->dict.nMarkers; ip.maxCorr = 0; ip.maxBorderErr = -1;            // only the extracted bits are wanted: no dictionary scan
    ip.detectInverted = 0; ip.minOtsuStdDev = d->prm.minOtsuStdDev; ip.W = W; ip.H = H; ip.pitch = gpitch; ip.frame_stride = gpitch * H; ip.max_cand = d->max_cand;
    ip.marks = nullptr;
    ip.pyr.n = 0;
    unsigned long long *d_codes = d->d_idcodes;
    ip.codes = d_codes;
    const int S = (ip.markerSize + 2 * ip.borderBits) * ip.cellSize;
    k_homography<<<dim3(4, 1), 64, 0, st>>>(fa, d->d_wM, S, d->max_cand, ip.pyr);
    k_identify<false><<<dim3(48, 1), ID_THREADS, identify_smem_bytes(S), st>>>(gray, d->d_dict, d->d_wM, fa, ip);
    codes.resize((size_t)nw);
    CU(cudaMemcpyAsync(codes.data(), d_codes, (size_t)nw * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    d->last_call_batch = 0;
    return launch_err("refine: bit extraction");
}

extern "C" int b2a_refine_detected_markers(b2a_detector *d, const b2a_frames *image, const b2a_board *board, float *corners, int32_t *ids, int *n_detected,
                                           int capacity, float *rejected, int *n_rejected, const b2a_camera *cam, const b2a_refine_params *rp_in,
                                           int32_t *recovered_idx, int *n_recovered)
{
    if (!d || !image || !board || !corners || !ids || !n_detected || !rejected || !n_rejected) return set_err(B2A_ERR_INVALID, "null argument");
    if (n_recovered) *n_recovered = 0;
    b2a_refine_params rp;
    if (rp_in) rp = *rp_in; else b2a_default_refine_params(&rp);
    if (!(rp.minRepDistance > 0)) return set_err(B2A_ERR_INVALID, "minRepDistance must be positive");            // CV_Assert in refineDetectedMarkers
    if (board->n_markers <= 0 || !board->ids || !board->obj_points) return set_err(B2A_ERR_INVALID, "empty board");
    if (image->batch != 1) return set_err(B2A_ERR_INVALID, "refineDetectedMarkers takes one image");
    if (d->prm.useAruco3Detection) return set_err(B2A_ERR_UNSUPPORTED, "refineDetectedMarkers on an ArUco3 detector (its rejected candidates are in the coordinates of the reduced image)");
    TRY(check_frames(d, image));
    const int nd = *n_detected, nr = *n_rejected, nbm = board->n_markers;
    if (nd < 0 || nr < 0 || nd > capacity) return set_err(B2A_ERR_INVALID, "bad counts");
    if (nd == 0 || nr == 0) return B2A_OK;
    CU(cudaSetDevice(d->device));
    // ---- where the undetected markers of the board should be (board_core.h) ----
    std::vector<int> und;                                   // board entries without a detection, board order
    std::vector<float> und_c;                               // their projected corners, Point2f like cv2's vectors
    auto detected = [&](int id) { for (int i = 0; i < nd; ++i) if (ids[i] == id) return i; return -1; };
    if (!cam) {
        const float z0 = board->obj_points[2];
        for (int i = 0; i < nbm * 4; ++i) if (board->obj_points[3 * i + 2] != z0) return set_err(B2A_ERR_INVALID, "board points must share one z for the homography form");
        std::vector<double> src, dst;
        for (int j = 0; j < nbm; ++j) {
            const int i = detected(board->ids[j]);
            if (i < 0) { und.push_back(j); continue; }
            for (int c = 0; c < 4; ++c) {
                src.push_back(board->obj_points[(j * 4 + c) * 3]); src.push_back(board->obj_points[(j * 4 + c) * 3 + 1]);
                dst.push_back(corners[i * 8 + 2 * c]); dst.push_back(corners[i * 8 + 2 * c + 1]);
            }
        }
        if (dst.empty() || und.empty()) return B2A_OK;
        double Hm[9];
        if (!homography_ls(src.data(), dst.data(), (int)dst.size() / 2, Hm, 10)) return set_err(B2A_ERR_INVALID, "degenerate homography");
        for (int j : und)
            for (int c = 0; c < 4; ++c) {
                double x, y;
                homography_apply(Hm, board->obj_points[(j * 4 + c) * 3], board->obj_points[(j * 4 + c) * 3 + 1], x, y);
                und_c.push_back((float)x); und_c.push_back((float)y);
            }
    } else {
        std::vector<double> obj, img;
        for (int i = 0; i < nd; ++i) {                      // Board::matchImagePoints: detection order
            int j = -1;
            for (int k = 0; k < nbm; ++k) if (board->ids[k] == ids[i]) { j = k; break; }
            if (j < 0) continue;
            for (int c = 0; c < 4; ++c) {
                for (int k = 0; k < 3; ++k) obj.push_back(board->obj_points[(j * 4 + c) * 3 + k]);
                img.push_back(corners[i * 8 + 2 * c]); img.push_back(corners[i * 8 + 2 * c + 1]);
            }
        }
        if (img.size() / 2 < 4) return B2A_OK;
        for (int j = 0; j < nbm; ++j) if (detected(board->ids[j]) < 0) und.push_back(j);
        if (und.empty()) return B2A_OK;
        double rvec[3], tvec[3], R[9];
        const Camera c = to_camera(cam);
        const int rc = board_pose(c, obj.data(), img.data(), (int)img.size() / 2, rvec, tvec);
        if (rc == 2) return set_err(B2A_ERR_INVALID, "board corners in general position need at least 6 matched points (cv2: DLT algorithm needs at least 6 points)");
        if (rc) return set_err(B2A_ERR_INVALID, "degenerate board pose");
        rodrigues_to_R(rvec, R);
        for (int j : und)
            for (int k = 0; k < 4; ++k) {
                const double X[3] = {board->obj_points[(j * 4 + k) * 3], board->obj_points[(j * 4 + k) * 3 + 1], board->obj_points[(j * 4 + k) * 3 + 2]};
                double u, v;
                project_point(c, R, tvec, X, u, v);
                und_c.push_back((float)u); und_c.push_back((float)v);
            }
    }
    // ---- bits of every candidate in its four corner orders (GPU) ----
    std::vector<unsigned long long> codes;
    if (rp.errorCorrectionRate >= 0 || d->prm.cornerRefinementMethod == 1) TRY(refine_extract_codes(d, image, rejected, nr, codes));   // also brings the gray frame to the device
    const int maxCorr = (int)((double)d->dict.maxCorrectionBits * (double)rp.errorCorrectionRate);
    auto dict_code = [&](int id) {
        unsigned long long v = 0;
        for (int k = 0; k < d->dict.nBytes; ++k) v |= (unsigned long long)d->dict_bytes[((size_t)id * 4) * d->dict.nBytes + k] << (8 * k);
        return v;
    };
    // ---- the greedy loop of refineDetectedMarkers ----
    std::vector<char> taken((size_t)nr, 0);
    std::vector<int> rec;
    int n_out = nd;
    for (size_t u = 0; u < und.size(); ++u) {
        const int uid = board->ids[und[u]];
        const float *uc = und_c.data() + u * 8;
        int bestj = -1, bestrot = 0;
        double bestd = (double)rp.minRepDistance * (double)rp.minRepDistance + 1;
        for (int j = 0; j < nr; ++j) {
            if (taken[j]) continue;
            bool valid = false;
            int rot = 0;
            double mind = bestd + 1;
            for (int c = 0; c < 4; ++c) {                    // the last corner order that beats the best so far stays (cv2 keeps overwriting)
                double cur = 0;
                for (int k = 0; k < 4; ++k) {
                    const float dx = uc[2 * k] - rejected[j * 8 + 2 * ((c + k) & 3)], dy = uc[2 * k + 1] - rejected[j * 8 + 2 * ((c + k) & 3) + 1];
                    const double dd = (double)(dx * dx + dy * dy);
                    if (dd > cur) cur = dd;
                }
                if (cur < bestd) { valid = true; rot = c; mind = cur; }
                if (!rp.checkAllOrders) break;
            }
            if (!valid) continue;
            int dist = 0;
            if (rp.errorCorrectionRate >= 0) {
                if (uid < 0 || uid >= d->dict.nMarkers) return set_err(B2A_ERR_INVALID, "board id outside the dictionary");
                dist = popc64(codes[(size_t)j * 4 + (rp.checkAllOrders ? rot : 0)] ^ dict_code(uid));
            }
            if (rp.errorCorrectionRate < 0 || dist < maxCorr) { bestj = j; bestd = mind; bestrot = rp.checkAllOrders ? rot : 0; }
        }
        if (bestj < 0) continue;
        if (n_out >= capacity) return set_err(B2A_ERR_CAPACITY, "detected markers exceed the caller's capacity");
        taken[bestj] = 1;
        for (int k = 0; k < 4; ++k) { corners[n_out * 8 + 2 * k] = rejected[bestj * 8 + 2 * ((bestrot + k) & 3)]; corners[n_out * 8 + 2 * k + 1] = rejected[bestj * 8 + 2 * ((bestrot + k) & 3) + 1]; }
        ids[n_out++] = uid;
        rec.push_back(bestj);
    }
    if (rec.empty()) return B2A_OK;
    // recovered corners get the detector's sub-pixel refinement, like accepted ones
    if (d->prm.cornerRefinementMethod == 1) {
        const int m = (int)rec.size();
        cudaStream_t st = d->stream;
        const int32_t cnt = m;
        CU(cudaMemcpyAsync(d->fo0.corners, corners + (size_t)nd * 8, (size_t)m * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d->fo0.n_accepted, &cnt, sizeof(cnt), cudaMemcpyHostToDevice, st));
        SubpixParams sp;
        const bool bgr = image->channels == 3;
        const size_t gp = bgr ? d->gray_pitch : (image->on_device ? (image->row_stride ? image->row_stride : (size_t)image->width) : (size_t)image->width);
        sp.W = image->width; sp.H = image->height; sp.pitch = gp; sp.frame_stride = gp * image->height; sp.max_markers = d->max_markers;
        sp.markerSize = d->dict.markerSize; sp.borderBits = d->prm.markerBorderBits; sp.maxWin = d->prm.cornerRefinementWinSize;
        sp.maxIter = d->prm.cornerRefinementMaxIterations; sp.relWin = d->prm.relativeCornerRefinmentWinSize; sp.eps = d->prm.cornerRefinementMinAccuracy;
        sp.fixedWin = 0; sp.refine = 1; sp.mul0 = 1.f; sp.mul1 = 1.f;
        const uint8_t *gray = bgr ? d->d_gray : (image->on_device ? image->data : d->d_in);
        if (m > d->max_markers) return set_err(B2A_ERR_CAPACITY, "recovered markers exceed max_markers");
        launch_subpix(d, st, gray, d->fo0.n_accepted, d->fo0.corners, d->d_corners2, 1, sp);
        CU(cudaMemcpyAsync(corners + (size_t)nd * 8, d->d_corners2, (size_t)m * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        TRY(launch_err("refine: k_subpix"));
    }
    // rejected list without the recovered candidates, order kept
    int w = 0;
    for (int j = 0; j < nr; ++j) if (!taken[j]) { if (w != j) std::memmove(rejected + (size_t)w * 8, rejected + (size_t)j * 8, 8 * sizeof(float)); ++w; }
    *n_rejected = w; *n_detected = n_out;
    for (size_t k = 0; k < rec.size(); ++k) if (recovered_idx) recovered_idx[k] = rec[k];
    if (n_recovered) *n_recovered = (int)rec.size();
    return B2A_OK;
}

extern "C" int b2a_estimate_pose_single_markers(b2a_detector *d, const float *corners, int n, const b2a_camera *cam, double *rvecs, double *tvecs)
{
    if (!d || !cam || (n > 0 && (!corners || !rvecs || !tvecs))) return set_err(B2A_ERR_INVALID, "null argument");
    if (!(cam->marker_length > 0)) return set_err(B2A_ERR_INVALID, "markerLength <= 0");
    if (n <= 0) return B2A_OK;
    if (d->in_flight) return set_err(B2A_ERR_INVALID, "a submitted batch is still in flight on this handle");
    CU(cudaSetDevice(d->device));
    // the handle's own output arrays serve as scratch (no allocation per call); more markers than they hold go in chunks
    const int cap = d->cfg.max_batch * d->max_markers;
    cudaStream_t st = d->stream;
    for (int o = 0; o < n; o += cap) {
        const int m = std::min(cap, n - o);
        CU(cudaMemcpyAsync(d->d_corners2, corners + (size_t)o * 8, (size_t)m * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
        k_pose<<<(m * 32 + POSE_THREADS - 1) / POSE_THREADS, POSE_THREADS, 0, st>>>(d->d_corners2, nullptr, 1, m, to_camera(cam), cam->marker_length, d->d_rvecs, d->d_tvecs, nullptr);
        CU(cudaMemcpyAsync(rvecs + (size_t)o * 3, d->d_rvecs, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(tvecs + (size_t)o * 3, d->d_tvecs, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    d->last_call_batch = 0;                                   // the result arrays no longer hold a detect call's poses
    return launch_err("k_pose");
}

// ------------------------------------------------------------------------------------------------
// stage taps
// ------------------------------------------------------------------------------------------------
extern "C" int b2a_debug_threshold(b2a_detector *d, const b2a_frames *f, uint8_t *gray, uint8_t *masks)
{
    if (d && d->prm.useAruco3Detection) return set_err(B2A_ERR_UNSUPPORTED, "stage taps of an ArUco3 detector (its front end runs on the reduced image)");
    std::vector<Sub> subs;
    TRY(run_pipeline(d, f, nullptr, 1, 0, &subs));
    const DetGeom g = make_geom(d, f->width, f->height, f->batch);
    cudaStream_t st = d->stream;
    if (gray)
        for (const Sub &s : subs)
            for (int b = 0; b < s.nb; ++b)
                CU(cudaMemcpy2DAsync(gray + (size_t)(s.b0 + b) * g.W * g.H, g.W, s.gray + (size_t)b * s.frame_stride, s.pitch, g.W, g.H, cudaMemcpyDeviceToHost, st));
    if (masks) {
        uint8_t *tmp = nullptr;
        const size_t n = (size_t)g.B * g.nScales * g.H * g.W;
        CU(cudaMalloc(&tmp, n));
        k_unpack_masks<<<d->num_sms * 8, 256, 0, st>>>(d->d_masks, tmp, g);
        cudaMemcpyAsync(masks, tmp, n, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(tmp);
        if (e != cudaSuccess) return set_err(B2A_ERR_CUDA, cudaGetErrorString(e));
    }
    CU(cudaStreamSynchronize(st));
    return launch_err("debug_threshold");
}

extern "C" int b2a_debug_contours(b2a_detector *d, const b2a_frames *f, int32_t *counts, int32_t *n_kept, int32_t *kept_len, int cap, int16_t *pts, int pts_cap)
{
    if (!f) return set_err(B2A_ERR_INVALID, "null argument");
    if (d && d->prm.useAruco3Detection) return set_err(B2A_ERR_UNSUPPORTED, "stage taps of an ArUco3 detector (its front end runs on the reduced image)");
    TRY(run_pipeline(d, f, nullptr, 1, 2 * f->width * f->height + 16, nullptr));     // exact total count: never give up on long borders
    const DetGeom g = make_geom(d, f->width, f->height, f->batch);
    const size_t FS = (size_t)g.B * g.nScales;
    std::vector<int> cc(FS), iso(FS), sc(FS), off((size_t)FS * g.surv_cap);
    std::vector<uint4> sorted((size_t)FS * g.surv_cap);
    CU(cudaMemcpy(cc.data(), d->d_contour_count, FS * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(iso.data(), d->d_iso_count, FS * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sc.data(), d->d_surv_count, FS * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(off.data(), d->d_pts_off, off.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sorted.data(), d->d_sorted, sorted.size() * sizeof(uint4), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> p((size_t)g.pts_cap);
    for (size_t fs = 0; fs < FS; ++fs) {
        if (counts) counts[fs] = cc[fs] + iso[fs];
        if (n_kept) n_kept[fs] = sc[fs];
        if (pts) CU(cudaMemcpy(p.data(), d->d_pts + fs * g.pts_cap, (size_t)g.pts_cap * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        int w = 0;
        for (int i = 0; i < sc[fs]; ++i) {
            const uint4 e = sorted[fs * g.surv_cap + i];
            if (kept_len && i < cap) kept_len[fs * cap + i] = (int)e.y;
            const int o = off[fs * g.surv_cap + i];
            if (pts && o >= 0)
                for (int k = 0; k < (int)e.y && w < pts_cap; ++k, ++w) {
                    pts[(fs * pts_cap + w) * 2] = (int16_t)(p[o + k] & 0xFFFF);
                    pts[(fs * pts_cap + w) * 2 + 1] = (int16_t)(p[o + k] >> 16);
                }
        }
    }
    return B2A_OK;
}

extern "C" int b2a_debug_candidates(b2a_detector *d, const b2a_frames *f, int32_t *n_cand, float *quads, int cap)
{
    if (d && d->prm.useAruco3Detection) return set_err(B2A_ERR_UNSUPPORTED, "stage taps of an ArUco3 detector (its front end runs on the reduced image)");
    TRY(run_pipeline(d, f, nullptr, 2, 0, nullptr));
    const int B = f->batch;
    std::vector<int> cnt((size_t)B * 8);
    CU(cudaMemcpy(cnt.data(), d->fs0.counters, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) {
        const int n = cnt[(size_t)b * 8 + FC_NCAND];
        if (n_cand) n_cand[b] = n;
        if (quads && n > 0) CU(cudaMemcpy(quads + (size_t)b * cap * 8, d->fs0.cq + (size_t)b * d->max_cand * 8, (size_t)std::min(n, cap) * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    return B2A_OK;
}

// ------------------------------------------------------------------------------------------------
// SLAM handle: observation mapping + EKF with Sigma resident on the device
// ------------------------------------------------------------------------------------------------
struct b2a_slam {
    b2a_slam_params p;
    int device = 0;
    cudaStream_t stream = nullptr;
    int N = 3, LD = 0, cap_lm = 0;
    double *d_sigma2 = nullptr;                      // ping-pong partner of d_sigma (cooperative kernel)
    double *d_mu = nullptr, *d_mus = nullptr, *d_sigma = nullptr, *d_K = nullptr, *d_GS = nullptr, *d_scratch = nullptr;
    int coop_grid = 0;                               // co-resident CTAs of k_ekf_frame (0 = use the per-observation kernels)
    bool panel = true;                               // all corrections of a frame as one rank-3M update (B2A_EKF_PANEL=0: one pass per observation)
    double *d_pze = nullptr, *d_pU0 = nullptr, *d_pK = nullptr, *d_pV0 = nullptr, *d_pV = nullptr, *d_pfac = nullptr, *d_pfacT = nullptr;
    // per-frame scratch, allocated once (grown only when a frame brings more markers than ever before)
    int obs_cap = 0;
    float *d_c = nullptr; int32_t *d_i = nullptr; double *d_r = nullptr, *d_t = nullptr;      // detections of the host-array entry point
    Observation *h_obs = nullptr; int *h_keep = nullptr, *h_n = nullptr;                      // pinned, written by k_observations; two sets
                                                                                              // (h_obs + set * obs_cap ...): one per frame in flight
    EkfObs *d_ekf = nullptr, *h_ekf[2] = {nullptr, nullptr}; cudaEvent_t ev_ekf[2] = {nullptr, nullptr}; int ekf_buf = 0;   // corrections of a frame
    // a frame whose landmarks are all known is one copy + four kernels of fixed shape per (N, corrections, staging buffer): replayed as a CUDA graph
    static constexpr int POSE_SLOTS = 8;
    double *h_pose = nullptr;                        // pinned [POSE_SLOTS][12]: mu[0..2], Sigma[0..2][0..2] of b2a_slam_robot_pose_submit
    cudaEvent_t ev_pose[POSE_SLOTS] = {};
    bool pose_pending[POSE_SLOTS] = {};
    struct EkfGraph { int N, nb, buf; cudaGraphExec_t exec; unsigned long long last_use; };
    std::vector<EkfGraph> ekf_graphs;
    unsigned long long ekf_graph_tick = 0;
    bool ekf_graph_ok = true;
    std::vector<int32_t> ids;                       // landmark k -> aruco id
    std::unordered_map<int32_t, int> id_index;      // aruco_id_map (aruco_slam.h:164): id -> landmark index, first insertion wins (:256)
    std::vector<int32_t> last_ids; std::vector<double> last_obs;   // last_observed_marker_ (NaN = unset)
    bool is_init = false;
};

extern "C" void b2a_default_slam_params(b2a_slam_params *p)
{
    p->Q_k = 0.01; p->R_x = 100.0; p->R_y = 100.0; p->R_theta = 10.0;       // parameters.yaml:5-8
    p->kl = 0.05; p->kr = 0.05; p->b = 0.09;                                 // parameters.yaml:11-13
    p->r2c_tx = 0.0; p->r2c_ty = 0.0;
    p->useful_distance_threshold = 3.f;                                      // aruco_slam.h:58 (the yaml key is never read)
    p->max_landmarks = 512;
}

static void slam_free_scratch(b2a_slam *s)
{
    for (auto &g : s->ekf_graphs) cudaGraphExecDestroy(g.exec);        // their nodes hold the addresses freed below
    s->ekf_graphs.clear();
    cudaFree(s->d_c); cudaFree(s->d_i); cudaFree(s->d_r); cudaFree(s->d_t); cudaFree(s->d_ekf);
    cudaFreeHost(s->h_obs); cudaFreeHost(s->h_keep); cudaFreeHost(s->h_ekf[0]); cudaFreeHost(s->h_ekf[1]);
    s->d_c = nullptr; s->d_i = nullptr; s->d_r = s->d_t = nullptr; s->d_ekf = nullptr; s->h_obs = nullptr; s->h_keep = nullptr; s->h_ekf[0] = s->h_ekf[1] = nullptr;
    s->obs_cap = 0;
}

// scratch for n markers per frame; called with the stream idle (only grows, from 256)
static int slam_reserve(b2a_slam *s, int n)
{
    if (n <= s->obs_cap) return B2A_OK;
    CU(cudaStreamSynchronize(s->stream));
    slam_free_scratch(s);
    const size_t c = (size_t)std::max(256, 2 * n);
    if (cudaMalloc(&s->d_c, c * 32) || cudaMalloc(&s->d_i, c * 4) || cudaMalloc(&s->d_r, c * 24) || cudaMalloc(&s->d_t, c * 24) ||
        cudaMalloc(&s->d_ekf, c * sizeof(EkfObs)) || cudaMallocHost(&s->h_obs, b2a_detector::MAX_CTX * c * sizeof(Observation)) || cudaMallocHost(&s->h_keep, b2a_detector::MAX_CTX * c * 4) ||
        cudaMallocHost(&s->h_ekf[0], c * sizeof(EkfObs)) || cudaMallocHost(&s->h_ekf[1], c * sizeof(EkfObs))) {
        slam_free_scratch(s);
        (void)cudaGetLastError();
        return set_err(B2A_ERR_CUDA, "cudaMalloc (SLAM frame scratch)");
    }
    s->obs_cap = (int)c;
    return B2A_OK;
}

extern "C" void b2a_slam_destroy(b2a_slam *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    slam_free_scratch(s);
    cudaFreeHost(s->h_n);
    for (cudaEvent_t e : s->ev_ekf) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s->ev_pose) if (e) cudaEventDestroy(e);
    cudaFreeHost(s->h_pose);
    cudaFree(s->d_sigma2);
    cudaFree(s->d_pze); cudaFree(s->d_pU0); cudaFree(s->d_pK); cudaFree(s->d_pV0); cudaFree(s->d_pV); cudaFree(s->d_pfac); cudaFree(s->d_pfacT);
    cudaFree(s->d_mu); cudaFree(s->d_mus); cudaFree(s->d_sigma); cudaFree(s->d_K); cudaFree(s->d_GS); cudaFree(s->d_scratch);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int b2a_slam_create(int device, const b2a_slam_params *p, b2a_slam **out)
{
    if (!out) return set_err(B2A_ERR_INVALID, "null output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return set_err(B2A_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return set_err(B2A_ERR_INVALID, "bad device ordinal");
    b2a_slam *s = new b2a_slam();
    if (p) s->p = *p; else b2a_default_slam_params(&s->p);
    s->device = device;
    s->cap_lm = s->p.max_landmarks > 0 ? s->p.max_landmarks : 512;
    s->LD = ((3 + 3 * s->cap_lm) + 15) & ~15;
    auto fail = [&](const char *m) { std::string e = m; b2a_slam_destroy(s); return set_err(B2A_ERR_CUDA, e); };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    const size_t LD = s->LD;
    if (cudaMalloc(&s->d_mu, LD * 8) || cudaMalloc(&s->d_mus, LD * 8) || cudaMalloc(&s->d_sigma, LD * LD * 8) ||
        cudaMalloc(&s->d_K, LD * 3 * 8) || cudaMalloc(&s->d_GS, LD * 3 * 8) || cudaMalloc(&s->d_scratch, LD * 6 * 8))
        return fail("cudaMalloc (EKF state)");
    {   // one cooperative launch per frame needs all CTAs resident at once
        int coop = 0, per_sm = 0, sms = 0;
        const size_t smem = 3 * LD * sizeof(double);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (coop && cudaFuncSetAttribute(k_ekf_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ekf_frame, 512, smem) == cudaSuccess && per_sm > 0 && !std::getenv("B2A_EKF_PER_OBS"))
            s->coop_grid = sms * std::min(per_sm, 2);
        if (s->coop_grid > 0 && cudaMalloc(&s->d_sigma2, LD * LD * 8) != cudaSuccess) { s->coop_grid = 0; s->d_sigma2 = nullptr; }
        if (s->d_sigma2) cudaMemsetAsync(s->d_sigma2, 0, LD * LD * 8, s->stream);
        (void)cudaGetLastError();
    }
    {   // panel form: Jacobians, innovations, U0 / K (N x 96), V0 / V (96 x N), the 96 x 96 factors
        if (const char *e = std::getenv("B2A_EKF_PANEL")) s->panel = std::atoi(e) != 0;
        if (s->panel) {
            if (cudaMalloc(&s->d_pze, EP_K * 8) || cudaMalloc(&s->d_pU0, LD * EP_K * 8) || cudaMalloc(&s->d_pK, LD * EP_K * 8) ||
                cudaMalloc(&s->d_pV0, LD * EP_K * 8) || cudaMalloc(&s->d_pV, LD * EP_K * 8) || cudaMalloc(&s->d_pfac, EP_K * EP_K * 8) || cudaMalloc(&s->d_pfacT, EP_K * EP_K * 8))
                return fail("cudaMalloc (EKF panel)");
            if (cudaFuncSetAttribute(k_ekf_panel_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EP_SSMEM) ||
                cudaFuncSetAttribute(k_ekf_panel_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EG_SMEM))
                return fail("shared memory of the EKF panel kernels");
        }
    }
    cudaMemsetAsync(s->d_mu, 0, LD * 8, s->stream);
    cudaMemsetAsync(s->d_sigma, 0, LD * LD * 8, s->stream);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return fail("memset");
    if (cudaMallocHost(&s->h_n, b2a_detector::MAX_CTX * sizeof(int)) || cudaEventCreateWithFlags(&s->ev_ekf[0], cudaEventDisableTiming) ||
        cudaEventCreateWithFlags(&s->ev_ekf[1], cudaEventDisableTiming) || slam_reserve(s, 256) != B2A_OK)
        return fail("frame scratch");
    *out = s;
    return B2A_OK;
}

extern "C" int b2a_slam_dim(const b2a_slam *s) { return s ? s->N : 0; }
extern "C" void *b2a_slam_stream(const b2a_slam *s) { return s ? (void *)s->stream : nullptr; }

extern "C" int b2a_slam_synchronize(b2a_slam *s)
{
    if (!s) return set_err(B2A_ERR_INVALID, "null handle");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    return B2A_OK;
}

// ------------------------------------------------------------------------------------------------
// wire / on-disk formats either side of the path (host only): landmark map text, pose + covariance record, map cubes
// ------------------------------------------------------------------------------------------------
extern "C" void b2a_quaternion_from_rpy(double roll, double pitch, double yaw, double q[4])
{   // tf2::Quaternion::setRPY (the library the reference calls at map_loader.cpp:90 and aruco_slam.cpp:273,387)
    const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
    const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
}

namespace {
// formatted extraction the way `std::istringstream >> int / double` behaves on one line: skip blanks, convert the longest
// prefix, and once one extraction has failed every later one fails too
struct LineCursor {
    const char *p, *end;
    bool failed = false;
    void skip() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p; }
    bool get_int(int &v)
    {
        if (failed) return false;
        skip();
        std::string tok(p, end);
        char *e = nullptr;
        const long r = std::strtol(tok.c_str(), &e, 10);
        if (e == tok.c_str()) { failed = true; return false; }
        v = (int)r; p += e - tok.c_str();
        return true;
    }
    bool get_double(double &v)
    {
        if (failed) return false;
        skip();
        std::string tok(p, end);
        char *e = nullptr;
        const double r = std::strtod(tok.c_str(), &e);
        if (e == tok.c_str()) { failed = true; return false; }
        v = r; p += e - tok.c_str();
        return true;
    }
};
}  // namespace

extern "C" int b2a_map_parse(const char *text, size_t len, b2a_map_marker *out, int cap, int *n_out)
{
    if (!text || !n_out || (cap > 0 && !out)) return set_err(B2A_ERR_INVALID, "null argument");
    int n = 0;
    const char *p = text, *end = text + len;
    while (p < end) {
        const char *eol = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        LineCursor c{p, eol};
        p = eol < end ? eol + 1 : end;
        c.skip();
        if (c.p >= c.end) continue;                                   // blank line (map_loader.cpp:26-30)
        if (*c.p == '#') continue;                                    // comment (:32-36)
        if (!std::isdigit((unsigned char)*c.p)) { *n_out = 0; return B2A_OK; }   // malformed: the whole map is dropped (:44-50)
        b2a_map_marker m;
        std::memset(&m, 0, sizeof(m));
        int id = 0;
        if (!(c.get_int(id) && c.get_double(m.length) && c.get_double(m.x) && c.get_double(m.y))) continue;   // (:52-58)
        m.id = id;
        if (!c.get_double(m.z)) m.z = 0;                              // (:60-64)
        if (!c.get_double(m.roll)) m.roll = 0;                        // 6th field (:65-69 zeroes yaw and leaves roll unset; 0 is the intent)
        if (!c.get_double(m.pitch)) m.pitch = 0;                      // (:70-74)
        if (!c.get_double(m.yaw)) m.yaw = 0;                          // 8th field (:75-79)
        b2a_quaternion_from_rpy(m.roll, m.pitch, m.yaw, m.q);         // addMarker (:86-94)
        if (n < cap) out[n] = m;
        ++n;
    }
    *n_out = n;
    if (n > cap) return set_err(B2A_ERR_CAPACITY, "map has more markers than the output array");
    return B2A_OK;
}

extern "C" int b2a_map_load(const char *path, b2a_map_marker *out, int cap, int *n_out)
{
    if (!path || !n_out) return set_err(B2A_ERR_INVALID, "null argument");
    std::FILE *f = std::fopen(path, "rb");
    if (!f) { *n_out = 0; return set_err(B2A_ERR_INVALID, std::string("cannot open map file ") + path); }
    std::string text;
    char buf[4096];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    std::fclose(f);
    return b2a_map_parse(text.data(), text.size(), out, cap, n_out);
}

extern "C" void b2a_pack_robot_pose(const double mu[3], const double S[9], b2a_pose_with_covariance *out)
{
    std::memset(out, 0, sizeof(*out));
    out->position[0] = mu[0]; out->position[1] = mu[1]; out->position[2] = 0.1;                    // aruco_slam.cpp:381-385
    b2a_quaternion_from_rpy(0, 0, mu[2], out->orientation);                                        // :388-390
    static const int at[9] = {0, 1, 5, 6, 7, 11, 30, 31, 35};                                      // :399-407
    for (int i = 0; i < 9; ++i) out->covariance[at[i]] = S[i];
}

extern "C" void b2a_pack_map_marker(int index, double marker_length, const double lm[3], b2a_map_marker *out)
{   // aruco_slam.cpp:266-281
    b2a_map_marker m;
    std::memset(&m, 0, sizeof(m));
    m.id = index; m.length = marker_length;
    m.x = lm[0]; m.y = lm[1]; m.z = 0.3;
    m.roll = 0; m.pitch = 1.5708; m.yaw = lm[2];
    b2a_quaternion_from_rpy(m.roll, m.pitch, m.yaw, m.q);
    *out = m;
}

namespace {
// tf2::Matrix3x3::getRotation (Shepperd's method as tf2 writes it): row-major R -> (x, y, z, w)
void tf2_get_rotation(const double *m, double *q)
{
    const double trace = m[0] + m[4] + m[8];
    if (trace > 0.0) {
        double s = std::sqrt(trace + 1.0);
        q[3] = s * 0.5;
        s = 0.5 / s;
        q[0] = (m[7] - m[5]) * s; q[1] = (m[2] - m[6]) * s; q[2] = (m[3] - m[1]) * s;
    } else {
        const int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = s * 0.5;
        s = 0.5 / s;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * s;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * s;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
    }
}
// tf2::Quaternion product a * b, (x, y, z, w)
void quat_mul(const double *a, const double *b, double *o)
{
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
}  // namespace

extern "C" int b2a_pack_detected_markers(const int32_t *ids, const double *rvecs, const double *tvecs, int n, double marker_length,
                                         float useful_distance_threshold, const double r2c_q[4], const double r2c_t[3],
                                         b2a_map_marker *out, int cap, int *n_out)
{   // aruco_slam.cpp:324-347
    if (n < 0 || !n_out || (n > 0 && (!ids || !rvecs || !tvecs)) || (cap > 0 && !out)) return set_err(B2A_ERR_INVALID, "null argument");
    const double qi[4] = {0, 0, 0, 1}, t0[3] = {0, 0, 0};
    const double *qt = r2c_q ? r2c_q : qi, *tt = r2c_t ? r2c_t : t0;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const double *rv = rvecs + 3 * i, *tv = tvecs + 3 * i;
        const float dist = (float)std::sqrt(tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2]);
        if (dist > useful_distance_threshold) continue;
        if (cnt < cap) {
            b2a_map_marker m;
            std::memset(&m, 0, sizeof(m));
            m.id = ids[i]; m.length = marker_length;
            double R[9], q[4];
            rodrigues_to_R(rv, R);
            tf2_get_rotation(R, q);
            // tf2::doTransform(pose, pose, r2c): position qt * (p, 0) * qt^-1 + t, orientation qt * q
            const double p[4] = {tv[0], tv[1], tv[2], 0.0}, qc[4] = {-qt[0], -qt[1], -qt[2], qt[3]};
            double a[4], r[4];
            quat_mul(qt, p, a); quat_mul(a, qc, r);
            m.x = r[0] + tt[0]; m.y = r[1] + tt[1]; m.z = r[2] + tt[2];
            quat_mul(qt, q, m.q);
            out[cnt] = m;
        }
        ++cnt;
    }
    *n_out = cnt;
    if (cnt > cap) return set_err(B2A_ERR_CAPACITY, "more detected markers than the output array");
    return B2A_OK;
}

extern "C" int b2a_slam_robot_pose(b2a_slam *s, b2a_pose_with_covariance *out)
{
    if (!s || !out) return set_err(B2A_ERR_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    double mu[3], S[9];
    CU(cudaMemcpy(mu, s->d_mu, sizeof(mu), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy2D(S, 3 * 8, s->d_sigma, (size_t)s->LD * 8, 3 * 8, 3, cudaMemcpyDeviceToHost));
    b2a_pack_robot_pose(mu, S, out);
    return B2A_OK;
}

// the same record without stalling the host thread: _submit enqueues the 96-byte read-back behind everything the filter has
// been given so far, _wait blocks only until that copy has landed
extern "C" int b2a_slam_robot_pose_submit(b2a_slam *s, int slot)
{
    if (!s || slot < 0 || slot >= b2a_slam::POSE_SLOTS) return set_err(B2A_ERR_INVALID, "pose slot outside 0 .. 7");
    CU(cudaSetDevice(s->device));
    if (!s->h_pose) {
        CU(cudaMallocHost(&s->h_pose, b2a_slam::POSE_SLOTS * 12 * sizeof(double)));
        for (auto &e : s->ev_pose) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    double *h = s->h_pose + 12 * slot;
    CU(cudaMemcpyAsync(h, s->d_mu, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpy2DAsync(h + 3, 3 * 8, s->d_sigma, (size_t)s->LD * 8, 3 * 8, 3, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev_pose[slot], s->stream));
    s->pose_pending[slot] = true;
    return B2A_OK;
}

extern "C" int b2a_slam_robot_pose_wait(b2a_slam *s, int slot, b2a_pose_with_covariance *out)
{
    if (!s || !out || slot < 0 || slot >= b2a_slam::POSE_SLOTS) return set_err(B2A_ERR_INVALID, "bad argument");
    if (!s->pose_pending[slot]) return set_err(B2A_ERR_INVALID, "no pose read-back submitted in this slot");
    CU(cudaSetDevice(s->device));
    CU(cudaEventSynchronize(s->ev_pose[slot]));
    s->pose_pending[slot] = false;
    b2a_pack_robot_pose(s->h_pose + 12 * slot, s->h_pose + 12 * slot + 3, out);
    return B2A_OK;
}

extern "C" int b2a_slam_detected_map(b2a_slam *s, double marker_length, b2a_map_marker *out, int cap, int *n_out)
{
    if (!s || !n_out || (cap > 0 && !out)) return set_err(B2A_ERR_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    const int n = (s->N - 3) / 3;
    std::vector<double> mu((size_t)s->N);
    CU(cudaMemcpy(mu.data(), s->d_mu, (size_t)s->N * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; ++i) b2a_pack_map_marker(i, marker_length, &mu[3 + 3 * i], &out[i]);
    *n_out = n;
    if (n > cap) return set_err(B2A_ERR_CAPACITY, "more landmarks than the output array");
    return B2A_OK;
}

extern "C" int b2a_slam_get_state(b2a_slam *s, double *mu, double *sigma, int32_t *ids)
{
    if (!s) return set_err(B2A_ERR_INVALID, "null handle");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    if (mu) CU(cudaMemcpy(mu, s->d_mu, (size_t)s->N * 8, cudaMemcpyDeviceToHost));
    if (sigma) CU(cudaMemcpy2D(sigma, (size_t)s->N * 8, s->d_sigma, (size_t)s->LD * 8, (size_t)s->N * 8, s->N, cudaMemcpyDeviceToHost));
    if (ids) std::memcpy(ids, s->ids.data(), s->ids.size() * sizeof(int32_t));
    return B2A_OK;
}

extern "C" int b2a_slam_set_state(b2a_slam *s, int N, const double *mu, const double *sigma, const int32_t *ids)
{
    if (!s || !mu || !sigma) return set_err(B2A_ERR_INVALID, "null argument");
    if (N < 3 || (N - 3) % 3 || (N - 3) / 3 > s->cap_lm) return set_err(B2A_ERR_INVALID, "state dimension");
    if (N > 3 && !ids) return set_err(B2A_ERR_INVALID, "ids required");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaMemset(s->d_sigma, 0, (size_t)s->LD * s->LD * 8));
    CU(cudaMemcpy(s->d_mu, mu, (size_t)N * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy2D(s->d_sigma, (size_t)s->LD * 8, sigma, (size_t)N * 8, (size_t)N * 8, N, cudaMemcpyHostToDevice));
    s->N = N;
    s->ids.assign(ids, ids + (N - 3) / 3);
    s->id_index.clear();
    for (int k = 0; k < (N - 3) / 3; ++k) s->id_index.insert({ids[k], k});
    s->last_ids.clear(); s->last_obs.clear();
    s->is_init = true;
    return B2A_OK;
}

extern "C" int b2a_slam_add_encoder(b2a_slam *s, double wl, double wr, double dt)
{
    if (!s) return set_err(B2A_ERR_INVALID, "null handle");
    CU(cudaSetDevice(s->device));
    if (!s->is_init) { s->is_init = true; return B2A_OK; }           // the first message only latches the clock (:24-29)
    k_ekf_predict<<<1, 256, 0, s->stream>>>(s->d_sigma, s->d_mu, s->N, s->LD, wl, wr, dt, s->p.kl, s->p.kr, s->p.b, s->p.Q_k, s->d_scratch);
    return launch_err("k_ekf_predict");
}

static ObsParams obs_params(const b2a_slam *s, const b2a_camera *cam)
{
    ObsParams op;
    op.R_x = s->p.R_x; op.R_y = s->p.R_y; op.R_theta = s->p.R_theta; op.marker_length = (double)cam->marker_length;
    op.r2c_tx = s->p.r2c_tx; op.r2c_ty = s->p.r2c_ty; op.useful_distance_threshold = s->p.useful_distance_threshold;
    return op;
}

// the observations k_observations left in the pinned arrays, gated ones dropped, detection order (:325-374)
static int collect_observations(const b2a_slam *s, int n, b2a_observation *out, int set = 0)
{
    int k = 0;
    const Observation *ho = s->h_obs + (size_t)set * s->obs_cap;
    const int *hk = s->h_keep + (size_t)set * s->obs_cap;
    for (int i = 0; i < n; ++i) {
        if (!hk[i]) continue;
        b2a_observation &o = out[k++];
        const Observation &h = ho[i];
        o.aruco_id = h.aruco_id; o.aruco_index = -1; o.x = h.x; o.y = h.y; o.theta = h.theta;
        std::memcpy(o.cov, h.cov, sizeof(o.cov));
    }
    return k;
}

extern "C" int b2a_slam_make_observations(b2a_slam *s, const float *corners, const int32_t *ids, const double *rvecs, const double *tvecs,
                                          int n, const b2a_camera *cam, b2a_observation *out, int *n_out)
{
    if (!s || !cam || !n_out || (n > 0 && (!corners || !ids || !rvecs || !tvecs || !out))) return set_err(B2A_ERR_INVALID, "null argument");
    *n_out = 0;
    if (n <= 0) return B2A_OK;
    CU(cudaSetDevice(s->device));
    TRY(slam_reserve(s, n));
    cudaStream_t st = s->stream;
    CU(cudaMemcpyAsync(s->d_c, corners, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_i, ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_r, rvecs, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_t, tvecs, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    k_observations<<<(n + 63) / 64, 64, 0, st>>>(s->d_c, s->d_i, s->d_r, s->d_t, n, nullptr, n, to_camera(cam), obs_params(s, cam), s->h_obs, s->h_keep, nullptr);
    CU(cudaStreamSynchronize(st));
    *n_out = collect_observations(s, n, out);
    return launch_err("k_observations");
}

extern "C" int b2a_slam_update(b2a_slam *s, const b2a_observation *obs, int n)
{
    if (!s || (n > 0 && !obs)) return set_err(B2A_ERR_INVALID, "null argument");
    CU(cudaSetDevice(s->device));
    if (n <= 0) { s->last_ids.clear(); s->last_obs.clear(); return B2A_OK; }
    TRY(slam_reserve(s, n));
    // checkLandmark (:423-435), then the reference's std::priority_queue<ArucoMarker> (aruco_slam.h:85-88,190): ascending
    // landmark index, new landmarks (-1) first.  The order among equal indices is whatever libstdc++'s heap gives the
    // reference (it is built with GCC); the same container with the same comparison and push order reproduces it -- pinned
    // against a run of the reference itself in tests/golden/slam_*.npz.
    struct Item {
        int index, seq;
        bool operator<(const Item &o) const { return index > o.index; }
    };
    std::priority_queue<Item> pq;
    int n_new = 0;
    for (int i = 0; i < n; ++i) {
        const auto it = s->id_index.find(obs[i].aruco_id);
        const int idx = it == s->id_index.end() ? -1 : it->second;
        n_new += idx < 0;
        pq.push({idx, i});
    }
    if ((int)s->ids.size() + n_new > s->cap_lm) return set_err(B2A_ERR_CAPACITY, "landmark capacity exceeded");
    cudaStream_t st = s->stream;
    CU(cudaMemcpyAsync(s->d_mus, s->d_mu, (size_t)s->N * 8, cudaMemcpyDeviceToDevice, st));      // mu snapshot (:88)
    std::vector<int32_t> new_last_ids(n);
    std::vector<double> new_last_obs((size_t)n * 3, NAN);
    // consecutive known-landmark corrections are staged in a pinned buffer (two of them, alternating by frame; a buffer is
    // rewritten only after the copy that last read it has completed) and run as one launch; nothing here waits for the GPU
    const int buf = s->ekf_buf;
    s->ekf_buf ^= 1;
    CU(cudaEventSynchronize(s->ev_ekf[buf]));
    EkfObs *stage = s->h_ekf[buf];
    int staged = 0, batch0 = 0;
    auto flush = [&]() -> int {
        const int nb = staged - batch0;
        if (nb <= 0) return B2A_OK;
        const int N = s->N;
        if (s->panel) {
            // one rank-3M update per (at most 32) corrections: Sigma is read and written once
            auto issue = [&]() -> int {
                CU(cudaMemcpyAsync(s->d_ekf + batch0, stage + batch0, (size_t)nb * sizeof(EkfObs), cudaMemcpyHostToDevice, st));
                for (int o0 = 0; o0 < nb; o0 += EP_MAX_OBS) {
                    EkfPanel p;
                    p.obs = s->d_ekf + batch0 + o0; p.M = std::min(EP_MAX_OBS, nb - o0);
                    p.ze = s->d_pze; p.U0 = s->d_pU0; p.Kall = s->d_pK; p.V0 = s->d_pV0; p.Vall = s->d_pV; p.fac = s->d_pfac; p.facT = s->d_pfacT;
                    const int tiles = (N + EG_T - 1) / EG_T, kdim = (3 * p.M + 3) & ~3;
                    const int gx = std::max((N + 127) / 128, (EP_K * EP_K + 127) / 128);          // the y == M slice also clears the padding of the 96 x 96 matrix
                    k_ekf_panel_gather<<<dim3(gx, p.M + 1), 128, 0, st>>>(s->d_sigma, s->d_mus, N, s->LD, p);
                    k_ekf_panel_factor<<<1, EP_MAX_OBS * EP_MAX_OBS, 0, st>>>(p);
                    const int rc = (N + EP_SW * EP_SV - 1) / (EP_SW * EP_SV);
                    k_ekf_panel_solve<<<2 * rc, 32 * EP_SW, EP_SSMEM, st>>>(s->d_mu, N, s->LD, p, rc);
                    k_ekf_panel_gemm<<<dim3(tiles, tiles), 512, EG_SMEM, st>>>(s->d_sigma, N, s->LD, p, kdim);
                }
                return B2A_OK;
            };
            // a frame whose landmarks are all known has one such flush of fixed shape per (N, corrections, staging buffer): captured once, replayed
            static const bool graph_env = !(std::getenv("B2A_GRAPH") && std::atoi(std::getenv("B2A_GRAPH")) == 0);
            bool done = false;
            if (graph_env && s->ekf_graph_ok && batch0 == 0 && nb <= EP_MAX_OBS) {
                b2a_slam::EkfGraph *ge = nullptr;
                for (auto &g : s->ekf_graphs) if (g.N == N && g.nb == nb && g.buf == buf) { ge = &g; break; }
                if (!ge) {
                    cudaGraph_t graph = nullptr;
                    cudaGraphExec_t exec = nullptr;
                    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
                    if (ce == cudaSuccess) {
                        const int rc = issue();
                        ce = cudaStreamEndCapture(st, &graph);
                        if (rc != B2A_OK && ce == cudaSuccess) ce = cudaErrorUnknown;
                    }
                    if (ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
                    if (graph) cudaGraphDestroy(graph);
                    if (ce != cudaSuccess || !exec) { (void)cudaGetLastError(); s->ekf_graph_ok = false; }      // plain launches from now on
                    else {
                        if (s->ekf_graphs.size() >= 48) {
                            size_t lru = 0;
                            for (size_t i = 1; i < s->ekf_graphs.size(); ++i) if (s->ekf_graphs[i].last_use < s->ekf_graphs[lru].last_use) lru = i;
                            cudaGraphExecDestroy(s->ekf_graphs[lru].exec);
                            s->ekf_graphs.erase(s->ekf_graphs.begin() + (long)lru);
                        }
                        s->ekf_graphs.push_back(b2a_slam::EkfGraph{N, nb, buf, exec, 0});
                        ge = &s->ekf_graphs.back();
                    }
                }
                if (ge) { ge->last_use = ++s->ekf_graph_tick; CU(cudaGraphLaunch(ge->exec, st)); done = true; }
            }
            if (!done) TRY(issue());
        } else if (s->coop_grid > 0 && (N + s->coop_grid - 1) / s->coop_grid <= 64) {
            CU(cudaMemcpyAsync(s->d_ekf + batch0, stage + batch0, (size_t)nb * sizeof(EkfObs), cudaMemcpyHostToDevice, st));
            int LD = s->LD, n_obs = nb, N_ = N;
            const EkfObs *obs_p = s->d_ekf + batch0;
            void *args[] = {&s->d_sigma, &s->d_sigma2, &s->d_mu, &s->d_mus, &N_, &LD, &obs_p, &n_obs};
            CU(cudaLaunchCooperativeKernel((void *)k_ekf_frame, dim3(s->coop_grid), dim3(512), args, 3 * (size_t)s->LD * sizeof(double), st));
            if (nb & 1) std::swap(s->d_sigma, s->d_sigma2);             // an odd number of ping-pongs ends in the other buffer
        } else {
            for (int k = batch0; k < staged; ++k) {
                const EkfObs &eo = stage[k];
                k_ekf_gain<<<(N + 255) / 256, 256, 0, st>>>(s->d_sigma, s->d_mus, N, s->LD, eo, s->d_K, s->d_GS);
                dim3 grid((N + 2 * EK_TX - 1) / (2 * EK_TX), (N + EK_ROWS - 1) / EK_ROWS);
                k_ekf_rank3<<<grid, dim3(EK_TX, EK_TY), 0, st>>>(s->d_sigma, s->d_mu, s->d_mus, N, s->LD, eo, s->d_K, s->d_GS);
            }
        }
        batch0 = staged;
        return B2A_OK;
    };
    for (int qi = 0; qi < n; ++qi) {
        const Item it = pq.top();
        pq.pop();
        const b2a_observation &o = obs[it.seq];
        EkfObs eo;
        eo.index = it.index;
        eo.z[0] = o.x; eo.z[1] = o.y; eo.z[2] = o.theta;
        std::memcpy(eo.Rk, o.cov, sizeof(eo.Rk));
        if (eo.index >= 0) {
            bool stationary = false;                                  // :192-198
            for (size_t l = 0; l < s->last_ids.size(); ++l)
                if (s->last_ids[l] == o.aruco_id) {
                    const double d0 = s->last_obs[3 * l] - o.x, d1 = s->last_obs[3 * l + 1] - o.y, d2 = s->last_obs[3 * l + 2] - o.theta;
                    if (std::sqrt(d0 * d0 + d1 * d1 + d2 * d2) < 0.01) stationary = true;
                    break;
                }
            if (!stationary) {
                new_last_obs[3 * qi] = o.x; new_last_obs[3 * qi + 1] = o.y; new_last_obs[3 * qi + 2] = o.theta;
                stage[staged++] = eo;
            }
        } else {
            TRY(flush());
            k_ekf_augment<<<1, 256, 0, st>>>(s->d_sigma, s->d_mu, s->d_mus, s->N, s->LD, eo);
            s->N += 3;
            s->id_index.insert({o.aruco_id, (int)s->ids.size()});      // :256 (insert: an id already in the map keeps its index)
            s->ids.push_back(o.aruco_id);
        }
        new_last_ids[qi] = o.aruco_id;
    }
    TRY(flush());
    CU(cudaEventRecord(s->ev_ekf[buf], st));
    s->last_ids.swap(new_last_ids); s->last_obs.swap(new_last_obs);    // :263
    return launch_err("EKF kernels");
}

static int add_image_checks(b2a_slam *s, b2a_detector *d, const b2a_frames *frame, const b2a_camera *cam)
{
    if (!s || !d || !frame || !cam) return set_err(B2A_ERR_INVALID, "null argument");
    if (frame->batch != 1) return set_err(B2A_ERR_INVALID, "add_image takes one frame");
    if (!(cam->marker_length > 0)) return set_err(B2A_ERR_INVALID, "markerLength <= 0");
    if (s->device != d->device) return set_err(B2A_ERR_INVALID, "detector and filter live on different devices");
    return B2A_OK;
}

// k_observations behind the detection of one frame: reads the context's device outputs, leaves the records in pinned set `set`
static int enqueue_observations(b2a_slam *s, b2a_detector *ctx, const b2a_camera *cam, int set, cudaStream_t st)
{
    const float *corners = ctx->prm.cornerRefinementMethod != 0 ? ctx->d_corners2 : ctx->fo0.corners;
    k_observations<<<(ctx->max_markers + 63) / 64, 64, 0, st>>>(corners, ctx->fo0.ids, ctx->d_rvecs, ctx->d_tvecs, 0, ctx->fo0.n_accepted, ctx->max_markers,
                                                                to_camera(cam), obs_params(s, cam), s->h_obs + (size_t)set * s->obs_cap,
                                                                s->h_keep + (size_t)set * s->obs_cap, s->h_n + set);
    ctx->launches++;
    return launch_err("k_observations");
}

extern "C" int b2a_slam_add_image(b2a_slam *s, b2a_detector *d, const b2a_frames *frame, const b2a_camera *cam)
{
    TRY(add_image_checks(s, d, frame, cam));
    if (!s->is_init) return B2A_OK;                                    // :84-85 (needs one encoder message first)
    TRY(slam_reserve(s, d->max_markers));
    // detect + pose + observation mapping in one enqueue: k_observations reads the detector's device outputs and leaves
    // the (few) observation records in pinned host memory; the call's one synchronisation covers it
    const std::function<int(cudaStream_t)> post = [&](cudaStream_t st) -> int { return enqueue_observations(s, d, cam, 0, st); };
    TRY(run_pipeline(d, frame, cam, 0, 0, nullptr, &post));
    if (d->h_status[0] != 0) return set_err(B2A_ERR_CAPACITY, "an internal list overflowed (see b2a_detections.status)");
    std::vector<b2a_observation> obs((size_t)std::max(s->h_n[0], 1));
    const int k = collect_observations(s, s->h_n[0], obs.data(), 0);
    return b2a_slam_update(s, obs.data(), k);
}

// The same split in two: the detection half of frame k + 1 can be enqueued (on the detector's other context) before the filter half of
// frame k runs, so a camera stream keeps two frames in flight.  State evolution is unchanged as long as the caller keeps the
// reference's order per frame: add_encoder (prediction) before the frame's wait (correction).
extern "C" int b2a_slam_add_image_submit(b2a_slam *s, b2a_detector *d, const b2a_frames *frame, const b2a_camera *cam, int *ticket)
{
    TRY(add_image_checks(s, d, frame, cam));
    if (!ticket) return set_err(B2A_ERR_INVALID, "null argument");
    *ticket = -1;
    if (!s->is_init) return B2A_OK;                                    // ignored like addImage before the first encoder message (:84-85)
    TRY(slam_reserve(s, d->max_markers));
    const int set = (int)(d->next_ticket % (unsigned)d->n_ctx);
    const std::function<int(b2a_detector *, cudaStream_t)> post = [&](b2a_detector *ctx, cudaStream_t st) -> int { return enqueue_observations(s, ctx, cam, set, st); };
    return submit_impl(d, frame, cam, ticket, &post);
}

extern "C" int b2a_slam_add_image_wait(b2a_slam *s, b2a_detector *d, int ticket)
{
    if (!s || !d) return set_err(B2A_ERR_INVALID, "null argument");
    if (ticket < 0) return B2A_OK;                                     // the submit was ignored
    b2a_detections det;
    const int rc = b2a_detect_pose_wait(d, ticket, &det);
    if (rc != B2A_OK) return rc;
    const int set = ticket % d->n_ctx;
    std::vector<b2a_observation> obs((size_t)std::max(s->h_n[set], 1));
    const int k = collect_observations(s, s->h_n[set], obs.data(), set);
    return b2a_slam_update(s, obs.data(), k);
}
