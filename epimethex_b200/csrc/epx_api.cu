/*
 * epx_api.cu — session state and the C ABI of libepx.so (include/epx.h).
 *
 * Host work done here is integer bookkeeping only (CSR validation, unique-probe slots, work items,
 * the Kolmogorov-Smirnov p-value lookup table, which depends on the group size alone); every
 * statistic of the path is computed by the kernels in epx_kernels.cu. There is no CPU fallback.
 */
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "epx_internal.h"
#include "epx_math.h"

using namespace epx;

namespace {

int fail(char *err, size_t errlen, int code, const char *fmt, ...)
{
    if (err && errlen) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define EPX_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(err, errlen, e_ == cudaErrorMemoryAllocation ? EPX_ENOMEM : EPX_ECUDA, "%s: %s", \
                        #call, cudaGetErrorString(e_));                                                  \
    } while (0)

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0; /* elements */
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

} // namespace

struct epx_session {
    int device = 0;
    int sm_count = 0;
    int host_threads = 8;    /* worker threads of the index build; epx_multi_create divides the host's cores among its sessions */
    cudaStream_t stream = nullptr;
    /* methylation, probe-major */
    DevBuf<double> meth_own;
    const double *meth = nullptr;
    int64_t P = 0, n_meth = 0, ld = 0;
    bool meth_partial = false; /* epx_analyze_host: only the rows the current index names are on the device */
    /* expression */
    DevBuf<double> expr;
    int64_t G = 0, n = 0;
    /* index */
    DevBuf<int32_t> probe_idx, pair_slot, uniq_probe, pair_gene;
    DevBuf<Item> items;
    int64_t Np = 0, U = 0, n_items = 0, G_index = -1;
    /* per-gene and per-probe intermediates */
    DevBuf<uint16_t> perm, ord;
    DevBuf<uint8_t> lab8;
    DevBuf<double> xc, gene_aux, lut, pt_tab, wt_tab;
    int pt_n = 0;
    int wt_m = -1, wt_N = 0, wt_ld = 0, wt_K = 0;   /* Welch (df, r) table of the pair kernel's t mode, built for group size wt_m */
    double wt_df0 = 0.0, wt_ddf = 1.0;
    DevBuf<int32_t> nuniq, scal; /* scal: counts[3], ref_gene, status */
    int64_t ldl = 0, ldx = 0, ldo = 0;
    int lut_m = -1, lut_n = -1;
    /* results */
    DevBuf<double> pair, pair_stat, gene;
    DevBuf<uint16_t> pair_na, gene_na;
    DevBuf<int32_t> pair_dnum, gene_dnum;
    /* timing */
    cudaEvent_t ev[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    cudaStream_t side = nullptr; /* gene-level kernels run here, concurrently with the probe ranking */
    bool ran = false, prepared = false;
    int launches = 0;
    PairParams pp;                      /* parameters of the pair kernel, valid after run_prepare */
    std::vector<int32_t> item_begin;    /* host copy: first pair of every work item, plus the total as a sentinel */
    void *staging = nullptr; /* pinned */
    size_t staging_bytes = 0;
    /* pooled probe groups (epx_session_run_groups) */
    const int32_t *h_probe_idx = nullptr, *h_pair_gene = nullptr; /* host copies of the index (in `staging`), to validate group lists */
    std::vector<int32_t> slot_of;                   /* [P] order row of a referenced methylation row (valid for the rows of the current index) */
    std::vector<uint8_t> ref_mark;                  /* [P] scratch of set_index: row is referenced; all 0 between calls */
    std::vector<int64_t> h_row_ptr;
    DevBuf<double> upload_stage;                    /* 2 x 64 MB: flat H2D chunks before re-spacing */
    cudaEvent_t ev_up[3] = { nullptr, nullptr, nullptr };
    bool join_side = false;                         /* run_pairs must join the side stream (gene tests) */
    int64_t index_P = -1;                           /* methylation row count the index was validated against */
    DevBuf<GroupDesc> g_desc;
    DevBuf<BlockDesc> g_blocks;
    DevBuf<TileDesc> g_tiles;
    DevBuf<int32_t> g_pairs, g_guard, g_multi;
    DevBuf<double> g_val[2], grp, grp_stat;
    DevBuf<uint8_t> g_lab[2];
    DevBuf<uint16_t> grp_na;
    DevBuf<int64_t> grp_dnum;
    int64_t n_groups = -1;
    DevBuf<unsigned long long> first_bad;           /* smallest row * n + sample of an uploaded value that is NaN / infinite */
};

struct epx_multi {
    std::vector<epx_session *> sess;
    /* probe-sliced mode (epx_multi_set_methylation_sliced): device i holds only the methylation rows
       rows_of[i][0..] (ascending), as its rows 0..; empty vectors = every device holds the whole matrix */
    std::vector<std::vector<int32_t>> rows_of;
    bool sliced = false;
    std::vector<int64_t> pair_lo;   /* after epx_multi_analyze: device i owns pairs [pair_lo[i], pair_lo[i+1]) */
    cudaEvent_t ev_expr = nullptr;  /* device 0: the expression matrix has arrived (the other devices copy it from there) */
};

/* contiguous gene ranges with (nearly) equal pair counts: cut[r] .. cut[r+1] is device r's range */
static std::vector<int64_t> gene_cuts(const int64_t *row_ptr, int64_t n_genes, int W)
{
    std::vector<int64_t> cut((size_t)W + 1, 0);
    cut[(size_t)W] = n_genes;
    const int64_t Np = row_ptr[n_genes];
    for (int r = 1; r < W; r++) {
        const int64_t target = Np * r / W;
        int64_t g = cut[(size_t)r - 1];
        while (g < n_genes && row_ptr[g] < target) g++;
        cut[(size_t)r] = g;
    }
    return cut;
}

/* run fn(i) for every device on its own thread; first failure wins */
template <typename F> static int for_each_device(epx_multi *m, F fn, char *err, size_t errlen)
{
    const size_t n = m->sess.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; i++)
        th.emplace_back([&, i]() {
            char buf[512] = "";
            rc[i] = fn((int)i, buf, sizeof buf);
            msg[i] = buf;
        });
    for (auto &t : th) t.join();
    for (size_t i = 0; i < n; i++)
        if (rc[i]) return fail(err, errlen, rc[i], "device %d: %s", m->sess[i]->device, msg[i].c_str());
    return EPX_OK;
}


extern "C" {

int epx_abi_version(void) { return EPX_ABI_VERSION; }

int epx_device_count(int *count, char *err, size_t errlen)
{
    if (!count) return fail(err, errlen, EPX_EINVAL, "count is NULL");
    EPX_CUDA(cudaGetDeviceCount(count));
    return EPX_OK;
}

int epx_device_info(int device, char *name, size_t namelen, int *sm_count, int64_t *mem_bytes, char *err,
                    size_t errlen)
{
    cudaDeviceProp prop;
    EPX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && namelen) snprintf(name, namelen, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (mem_bytes) *mem_bytes = (int64_t)prop.totalGlobalMem;
    return EPX_OK;
}

void epx_options_default(epx_options *opt)
{
    if (!opt) return;
    memset(opt, 0, sizeof(*opt));
    opt->struct_size = (int32_t)sizeof(*opt);
    opt->data_linear = 1;
    opt->test_expr_t = 1;
    opt->test_meth_t = 0;
}

/* page-locked for EVERY device of the process (portable): with the default flag only the device that was current at
 * the allocation sees it as pinned, and the other devices of an epx_multi handle copied their result rows into it at
 * the pageable rate (3.5 ms more per call on 2 GPUs) */
int epx_host_alloc(void **ptr, size_t bytes, char *err, size_t errlen)
{
    if (!ptr) return fail(err, errlen, EPX_EINVAL, "ptr is NULL");
    EPX_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return EPX_OK;
}

void epx_host_free(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int epx_session_create(int device, epx_session **out, char *err, size_t errlen)
{
    if (!out) return fail(err, errlen, EPX_EINVAL, "out is NULL");
    *out = nullptr;
    int count = 0;
    EPX_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) return fail(err, errlen, EPX_ECUDA, "no CUDA device: libepx has no CPU fallback");
    if (device < 0 || device >= count) return fail(err, errlen, EPX_EINVAL, "device %d out of range [0,%d)", device, count);
    EPX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    EPX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(err, errlen, EPX_ECUDA, "device %d is sm_%d%d; libepx is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    EPX_CUDA(kernels_init());
    epx_session *s = new (std::nothrow) epx_session();
    if (!s) return fail(err, errlen, EPX_ENOMEM, "out of host memory");
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&s->ev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking);
    for (int i = 0; i < 3 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = s->scal.reserve(16);
    if (e == cudaSuccess) e = cudaMemset(s->scal.p, 0, 16 * sizeof(int32_t)); /* [6], [7]: work queue of the pair kernel; [8], [9]: of the expression-row kernel; [10], [11]: of the probe-rank kernel; [12]-[13]: reference-gene maximum; [14]: rows sorted by the full network after a failed bucket placement (since the session was created) */
    if (e != cudaSuccess) {
        epx_session_destroy(s);
        return fail(err, errlen, EPX_ECUDA, "session setup: %s", cudaGetErrorString(e));
    }
    *out = s;
    return EPX_OK;
}

void epx_session_destroy(epx_session *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    s->meth_own.release(); s->expr.release(); s->probe_idx.release(); s->pair_slot.release();
    s->uniq_probe.release(); s->items.release(); s->pair_gene.release(); s->perm.release(); s->ord.release(); s->lab8.release();
    s->xc.release(); s->gene_aux.release(); s->lut.release(); s->pt_tab.release(); s->wt_tab.release(); s->nuniq.release(); s->scal.release();
    s->pair.release(); s->pair_stat.release(); s->gene.release(); s->pair_na.release(); s->gene_na.release();
    s->pair_dnum.release(); s->gene_dnum.release();
    s->g_desc.release(); s->g_blocks.release(); s->g_tiles.release(); s->g_pairs.release(); s->g_guard.release(); s->g_multi.release();
    s->g_val[0].release(); s->g_val[1].release(); s->g_lab[0].release(); s->g_lab[1].release();
    s->grp.release(); s->grp_stat.release(); s->grp_na.release(); s->grp_dnum.release(); s->first_bad.release();
    for (int i = 0; i < 8; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (int i = 0; i < 3; i++) if (s->ev_up[i]) cudaEventDestroy(s->ev_up[i]);
    s->upload_stage.release();
    if (s->side) cudaStreamDestroy(s->side);
    if (s->staging) cudaFreeHost(s->staging);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

static int check_samples(epx_session *s, int64_t n, char *err, size_t errlen)
{
    if (n < 3) return fail(err, errlen, EPX_EINVAL, "need at least 3 samples, got %lld", (long long)n);
    if (n > 16384) return fail(err, errlen, EPX_EINVAL, "n_samples %lld exceeds the supported maximum 16384", (long long)n);
    (void)s;
    return EPX_OK;
}

/* the upload kernels atomicMin the position of the first non-finite value into s->first_bad */
static cudaError_t arm_finite_check(epx_session *s, cudaStream_t st)
{
    cudaError_t e = s->first_bad.reserve(1);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(s->first_bad.p, 0xff, sizeof(unsigned long long), st);
}
/* after the stream has been synchronised: 0 when every value was finite, else EPX_EINVAL naming the first bad one */
static int finite_check_result(epx_session *s, const char *what, const char *unit, int64_t n, char *err, size_t errlen)
{
    unsigned long long bad = ~0ull;
    EPX_CUDA(cudaMemcpy(&bad, s->first_bad.p, sizeof bad, cudaMemcpyDeviceToHost));
    if (bad == ~0ull) return EPX_OK;
    return fail(err, errlen, EPX_EINVAL,
                "%s %s %lld holds a NaN or an infinite value (sample %lld): drop incomplete %ss before the call, as the "
                "reference does with na.omit (R/EpimethexAnalysis.R:114, :278)",
                what, unit, (long long)(bad / (unsigned long long)n), (long long)(bad % (unsigned long long)n), unit);
}

int epx_session_set_methylation_rows(epx_session *s, const double *rows, int64_t n_probes, int64_t n_samples,
                                     int64_t ld, char *err, size_t errlen)
{
    if (!s || !rows || n_probes < 1 || ld < n_samples) return fail(err, errlen, EPX_EINVAL, "bad methylation arguments");
    int rc = check_samples(s, n_samples, err, errlen);
    if (rc) return rc;
    EPX_CUDA(cudaSetDevice(s->device));
    const int64_t dld = round_up(n_samples, 4); /* 32-byte aligned rows */
    EPX_CUDA(s->meth_own.reserve((size_t)(n_probes * dld)));
    /* a failed upload leaves no usable matrix behind */
    s->meth = nullptr; s->meth_partial = false; s->prepared = false; s->ran = false; s->n_groups = -1;
    EPX_CUDA(arm_finite_check(s, s->stream));
    if (ld == dld) {
        EPX_CUDA(cudaMemcpyAsync(s->meth_own.p, rows, (size_t)(n_probes * dld) * 8, cudaMemcpyHostToDevice, s->stream));
        EPX_CUDA(launch_check_finite(s->meth_own.p, n_probes, n_samples, dld, s->first_bad.p, s->stream));
    } else if (ld == n_samples) {
        /* contiguous host matrix: flat copies at full PCIe rate into a staging buffer, rows re-spaced on the
         * device (a 2-D copy with 3.7 KB rows runs at a fraction of the link rate) */
        const int64_t chunk_rows = (64ll << 20) / (n_samples * 8) > 0 ? (64ll << 20) / (n_samples * 8) : 1;
        /* two staging buffers kept by the session: the re-spacing of chunk i (side stream) overlaps the copy of
         * chunk i+1 (main stream) */
        EPX_CUDA(s->upload_stage.reserve((size_t)(2 * chunk_rows * n_samples)));
        cudaEvent_t copied = s->ev_up[0], spaced[2] = { s->ev_up[1], s->ev_up[2] };
        cudaError_t e = cudaEventRecord(copied, s->stream); /* the side stream starts after the check word is armed */
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s->side, copied, 0);
        int64_t i = 0;
        for (int64_t r0 = 0; r0 < n_probes && e == cudaSuccess; r0 += chunk_rows, i++) {
            const int64_t nr = (n_probes - r0) < chunk_rows ? (n_probes - r0) : chunk_rows;
            double *buf = s->upload_stage.p + (size_t)((i & 1) * chunk_rows * n_samples);
            if (i >= 2) e = cudaStreamWaitEvent(s->stream, spaced[i & 1], 0); /* buffer free again */
            if (e == cudaSuccess) e = cudaMemcpyAsync(buf, rows + r0 * n_samples, (size_t)(nr * n_samples) * 8, cudaMemcpyHostToDevice, s->stream);
            if (e == cudaSuccess) e = cudaEventRecord(copied, s->stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s->side, copied, 0);
            if (e == cudaSuccess) e = launch_respace_rows(buf, s->meth_own.p + r0 * dld, nr, n_samples, dld, r0, s->first_bad.p, s->side);
            if (e == cudaSuccess) e = cudaEventRecord(spaced[i & 1], s->side);
        }
        cudaError_t e2 = cudaStreamSynchronize(s->side);
        if (e == cudaSuccess) e = e2;
        if (e != cudaSuccess) { cudaStreamSynchronize(s->stream); return fail(err, errlen, EPX_ECUDA, "H2D: %s", cudaGetErrorString(e)); }
    } else {
        EPX_CUDA(cudaMemcpy2DAsync(s->meth_own.p, (size_t)dld * 8, rows, (size_t)ld * 8, (size_t)n_samples * 8,
                                   (size_t)n_probes, cudaMemcpyHostToDevice, s->stream));
        EPX_CUDA(launch_check_finite(s->meth_own.p, n_probes, n_samples, dld, s->first_bad.p, s->stream));
    }
    EPX_CUDA(cudaStreamSynchronize(s->stream));
    rc = finite_check_result(s, "methylation", "probe", n_samples, err, errlen);
    if (rc) return rc;
    s->meth = s->meth_own.p; s->P = n_probes; s->n_meth = n_samples; s->ld = dld;
    /* the index (unique-probe slots, work items) depends on the row count only: it survives a new matrix of the
     * same shape; everything computed from the old values does not */
    if (s->index_P != n_probes) s->G_index = -1;
    s->prepared = false; s->ran = false; s->n_groups = -1;
    return EPX_OK;
}

int epx_session_set_methylation_cols(epx_session *s, const double *const *cols, int64_t n_samples, int64_t n_probes,
                                     char *err, size_t errlen)
{
    if (!s || !cols || n_probes < 1) return fail(err, errlen, EPX_EINVAL, "bad methylation arguments");
    int rc = check_samples(s, n_samples, err, errlen);
    if (rc) return rc;
    EPX_CUDA(cudaSetDevice(s->device));
    const int64_t dld = round_up(n_samples, 4);
    EPX_CUDA(s->meth_own.reserve((size_t)(n_probes * dld)));
    /* stage a chunk of sample columns on the device, transpose into probe-major rows */
    int chunk = (int)((int64_t)(256ll << 20) / (n_probes * 8));
    if (chunk < 1) chunk = 1;
    if (chunk > n_samples) chunk = (int)n_samples;
    DevBuf<double> stage;
    EPX_CUDA(stage.reserve((size_t)chunk * (size_t)n_probes));
    s->meth = nullptr; s->meth_partial = false; s->prepared = false; s->ran = false; s->n_groups = -1;
    EPX_CUDA(arm_finite_check(s, s->stream));
    for (int64_t s0 = 0; s0 < n_samples; s0 += chunk) {
        const int ns = (int)((n_samples - s0) < chunk ? (n_samples - s0) : chunk);
        for (int j = 0; j < ns; j++) {
            if (!cols[s0 + j]) { stage.release(); return fail(err, errlen, EPX_EINVAL, "column %lld is NULL", (long long)(s0 + j)); }
            cudaError_t e = cudaMemcpyAsync(stage.p + (size_t)j * n_probes, cols[s0 + j], (size_t)n_probes * 8,
                                            cudaMemcpyHostToDevice, s->stream);
            if (e != cudaSuccess) { stage.release(); return fail(err, errlen, EPX_ECUDA, "H2D: %s", cudaGetErrorString(e)); }
        }
        cudaError_t e = launch_transpose_cols(stage.p, s->meth_own.p, n_probes, ns, s0, dld, n_samples, s->first_bad.p, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) { stage.release(); return fail(err, errlen, EPX_ECUDA, "transpose: %s", cudaGetErrorString(e)); }
    }
    stage.release();
    rc = finite_check_result(s, "methylation", "probe", n_samples, err, errlen);
    if (rc) return rc;
    s->meth = s->meth_own.p; s->P = n_probes; s->n_meth = n_samples; s->ld = dld;
    /* the index (unique-probe slots, work items) depends on the row count only: it survives a new matrix of the
     * same shape; everything computed from the old values does not */
    if (s->index_P != n_probes) s->G_index = -1;
    s->prepared = false; s->ran = false; s->n_groups = -1;
    return EPX_OK;
}

int epx_session_set_methylation_device(epx_session *s, const double *d_rows, int64_t n_probes, int64_t n_samples,
                                       int64_t ld, char *err, size_t errlen)
{
    if (!s || !d_rows || n_probes < 1 || ld < n_samples) return fail(err, errlen, EPX_EINVAL, "bad methylation arguments");
    int rc = check_samples(s, n_samples, err, errlen);
    if (rc) return rc;
    if ((((uintptr_t)d_rows) & 15) != 0 || (ld & 1) != 0)
        return fail(err, errlen, EPX_EINVAL, "device methylation rows must be 16-byte aligned with an even leading dimension (bulk copies)");
    EPX_CUDA(cudaSetDevice(s->device));
    s->meth_own.release();
    s->meth = nullptr; s->meth_partial = false; s->prepared = false; s->ran = false; s->n_groups = -1;
    EPX_CUDA(arm_finite_check(s, s->stream));
    EPX_CUDA(launch_check_finite(d_rows, n_probes, n_samples, ld, s->first_bad.p, s->stream));
    EPX_CUDA(cudaStreamSynchronize(s->stream));
    rc = finite_check_result(s, "methylation", "probe", n_samples, err, errlen);
    if (rc) return rc;
    s->meth = d_rows; s->P = n_probes; s->n_meth = n_samples; s->ld = ld;
    /* the index (unique-probe slots, work items) depends on the row count only: it survives a new matrix of the
     * same shape; everything computed from the old values does not */
    if (s->index_P != n_probes) s->G_index = -1;
    s->prepared = false; s->ran = false; s->n_groups = -1;
    return EPX_OK;
}

/* the expression upload in two halves, so that epx_analyze can build the index on the host while the copy runs */
static int set_expression_begin(epx_session *s, const double *expr, int64_t n_genes, int64_t n_samples, int is_device,
                                char *err, size_t errlen)
{
    if (!s || !expr || n_genes < 1) return fail(err, errlen, EPX_EINVAL, "bad expression arguments");
    int rc = check_samples(s, n_samples, err, errlen);
    if (rc) return rc;
    EPX_CUDA(cudaSetDevice(s->device));
    /* results, intermediates and pooled groups of the previous expression matrix are void from here on: fetch /
     * result_device_ptr / run_pairs / run_groups answer EPX_ESTATE until the next run (their sizes follow G and n) */
    s->prepared = false; s->ran = false; s->n_groups = -1;
    s->G = 0; s->n = 0;
    EPX_CUDA(s->expr.reserve((size_t)(n_genes * n_samples)));
    EPX_CUDA(arm_finite_check(s, s->stream));
    EPX_CUDA(cudaMemcpyAsync(s->expr.p, expr, (size_t)(n_genes * n_samples) * 8,
                             is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s->stream));
    EPX_CUDA(launch_check_finite(s->expr.p, n_genes, n_samples, n_samples, s->first_bad.p, s->stream));
    return EPX_OK;
}

static int set_expression_finish(epx_session *s, int64_t n_genes, int64_t n_samples, char *err, size_t errlen)
{
    EPX_CUDA(cudaStreamSynchronize(s->stream));
    int rc = finite_check_result(s, "expression", "gene", n_samples, err, errlen);
    if (rc) return rc;
    s->G = n_genes; s->n = n_samples;
    return EPX_OK;
}

int epx_session_set_expression(epx_session *s, const double *expr, int64_t n_genes, int64_t n_samples, int is_device,
                               char *err, size_t errlen)
{
    int rc = set_expression_begin(s, expr, n_genes, n_samples, is_device, err, errlen);
    if (rc) return rc;
    return set_expression_finish(s, n_genes, n_samples, err, errlen);
}

/* host_call: epx_analyze_host builds the index BEFORE the rows it names are fetched (s->P is set, s->meth not yet) */
static int set_index_impl(epx_session *s, const int64_t *row_ptr, const int32_t *probe_idx, int64_t n_genes,
                          bool host_call, char *err, size_t errlen)
{
    if (!s || !row_ptr || n_genes < 1) return fail(err, errlen, EPX_EINVAL, "bad index arguments");
    if (!s->meth && !host_call) return fail(err, errlen, EPX_ESTATE, "set the methylation matrix before the index");
    if (row_ptr[0] != 0) return fail(err, errlen, EPX_EINVAL, "row_ptr[0] must be 0");
    const int64_t Np = row_ptr[n_genes];
    if (Np < 0 || Np > 0x7fffffff) return fail(err, errlen, EPX_EINVAL, "pair count %lld out of range", (long long)Np);
    if (Np > 0 && !probe_idx) return fail(err, errlen, EPX_EINVAL, "probe_idx is NULL");
    /* the same index again (re-run with other options, repeated calls on one gene range): everything derived
     * from it is still on the device */
    if (s->G_index == n_genes && s->Np == Np && s->h_row_ptr.size() == (size_t)n_genes + 1 && s->index_P == s->P &&
        memcmp(s->h_row_ptr.data(), row_ptr, ((size_t)n_genes + 1) * sizeof(int64_t)) == 0 &&
        (Np == 0 || memcmp(s->h_probe_idx, probe_idx, (size_t)Np * sizeof(int32_t)) == 0))
        return EPX_OK;
    if (s->meth_partial && !host_call)
        return fail(err, errlen, EPX_ESTATE,
                    "the session holds only the methylation rows of the last epx_analyze_host index: upload the matrix "
                    "(epx_session_set_methylation_*) before another index, or call epx_analyze_host again");
    EPX_CUDA(cudaSetDevice(s->device));

    /* Host bookkeeping of a new index. The reference is called once per gene range (README.md:90), so this runs on
     * every call: everything lives in buffers the session keeps -- one page-locked block for what goes to the device
     * (probe_idx | pair_slot | pair_gene | uniq | items: the copies are truly asynchronous and no fresh pages are
     * touched), a slot table that is reset entry by entry -- instead of five vectors allocated and faulted in per call. */
    const size_t max_items = (size_t)Np / 4 + (size_t)n_genes + 1;
    const size_t need = (size_t)Np * 4 * 4 + max_items * sizeof(Item) + 64;
    if (need > s->staging_bytes) {
        if (s->staging) cudaFreeHost(s->staging);
        s->staging = nullptr; s->staging_bytes = 0;
        EPX_CUDA(cudaHostAlloc(&s->staging, need + need / 4, cudaHostAllocDefault));
        s->staging_bytes = need + need / 4;
    }
    int32_t *h_pi = (int32_t *)s->staging, *pair_slot = h_pi + Np, *pair_gene = pair_slot + Np, *uniq = pair_gene + Np;
    Item *items = (Item *)(((uintptr_t)(uniq + Np) + 15) & ~(uintptr_t)15);
    /* work items of the pair kernel: up to ITEM_PAIRS consecutive pairs of one gene. Small jobs (a 1000-gene
     * range is ~23 k pairs for ~2400 resident warps) get smaller items so that every warp has several. */
    int64_t item_pairs = Np / ((int64_t)s->sm_count * 16 * 4);
    if (item_pairs < 4) item_pairs = 4;
    if (item_pairs > ITEM_PAIRS) item_pairs = ITEM_PAIRS;
    /* The slot (order row) of a referenced probe is its rank among the referenced probes, by row number: the rank kernel
     * then reads the matrix front to back, and the three passes below -- mark and copy, number the marked rows, look the
     * slots up -- split over host threads (first-appearance numbering was one serial pass of ~4 ms for 474,611 pairs,
     * on every call: the reference is called once per gene range, README.md:90). */
    if (s->slot_of.size() < (size_t)s->P) s->slot_of.assign((size_t)s->P, -1);
    if (s->ref_mark.size() < (size_t)s->P) s->ref_mark.assign((size_t)s->P, 0);
    int32_t *slot_of = s->slot_of.data();
    uint8_t *ref = s->ref_mark.data();
    const int64_t P = s->P;
    int nt = (int)std::thread::hardware_concurrency();
    if (nt > s->host_threads) nt = s->host_threads;
    if (nt < 1 || Np < 65536) nt = 1;
    std::vector<int64_t> cnt((size_t)nt + 1, 0), bad_at_t((size_t)nt, -1);
    std::vector<int> bad_code_t((size_t)nt, 0);
    std::atomic<int> arrived{0};
    auto barrier = [&](int phase) { /* all nt workers meet here; phases are numbered 1, 2, .. */
        arrived.fetch_add(1, std::memory_order_acq_rel);
        while (arrived.load(std::memory_order_acquire) < phase * nt) std::this_thread::yield();
    };
    auto work = [&](int tid) {
        /* pass 1: this worker's genes (cut at equal pair counts): checks, host copies, marks */
        const int64_t *gb = std::lower_bound(row_ptr, row_ptr + n_genes, Np * tid / nt);
        const int64_t *ge = tid + 1 == nt ? row_ptr + n_genes : std::lower_bound(row_ptr, row_ptr + n_genes, Np * (tid + 1) / nt);
        for (int64_t g = gb - row_ptr; g < ge - row_ptr && !bad_code_t[tid]; g++) {
            const int64_t b = row_ptr[g], e = row_ptr[g + 1];
            if (b < 0 || e < b || e > Np) { bad_code_t[tid] = 1; bad_at_t[tid] = g; break; }
            for (int64_t j = b; j < e; j++) {
                const int32_t pi = probe_idx[j];
                h_pi[j] = pi;
                pair_gene[j] = (int32_t)g;
                if (pi >= P) { bad_code_t[tid] = 2; bad_at_t[tid] = j; break; }
                if (pi >= 0) __atomic_store_n(&ref[pi], (uint8_t)1, __ATOMIC_RELAXED); /* several workers may store the same 1 */
            }
        }
        barrier(1);
        /* pass 2: number the marked rows of this worker's slice of the matrix */
        const int64_t p0 = P * tid / nt, p1 = P * (tid + 1) / nt;
        int64_t c = 0;
        for (int64_t q = p0; q < p1; q++) c += ref[q];
        cnt[(size_t)tid + 1] = c;
        barrier(2);
        int64_t base = 0;
        for (int k = 0; k <= tid; k++) base += cnt[(size_t)k];
        for (int64_t q = p0; q < p1; q++)
            if (ref[q]) { ref[q] = 0; slot_of[q] = (int32_t)base; uniq[base++] = (int32_t)q; }
        barrier(3);
        /* pass 3: the slot of every pair */
        const int64_t j0 = Np * tid / nt, j1 = Np * (tid + 1) / nt;
        for (int64_t j = j0; j < j1; j++) { const int32_t pi = h_pi[j]; pair_slot[j] = (pi < 0 || pi >= P) ? -1 : slot_of[pi]; }
    };
    {
        std::vector<std::thread> th;
        for (int k = 1; k < nt; k++) th.emplace_back(work, k);
        work(0);
        for (auto &t : th) t.join();
    }
    size_t n_uniq = 0, n_items = 0;
    for (int k = 1; k <= nt; k++) n_uniq += (size_t)cnt[(size_t)k];
    int bad_code = 0;
    int64_t bad_at = 0;
    for (int k = 0; k < nt && !bad_code; k++) if (bad_code_t[k]) { bad_code = bad_code_t[k]; bad_at = bad_at_t[k]; }
    if (!bad_code) { /* the genes must tile [0, Np) in order: the workers' slices rely on it */
        for (int64_t g = 0; g < n_genes; g++) if (row_ptr[g + 1] < row_ptr[g]) { bad_code = 1; bad_at = g; break; }
    }
    if (bad_code) {
        std::fill(s->ref_mark.begin(), s->ref_mark.end(), (uint8_t)0);
        s->G_index = -1;                                          /* the staging block no longer mirrors a valid index */
        if (bad_code == 1) return fail(err, errlen, EPX_EINVAL, "row_ptr is not monotone at gene %lld", (long long)bad_at);
        return fail(err, errlen, EPX_EINVAL, "probe_idx[%lld] = %d >= n_probes %lld", (long long)bad_at, probe_idx[bad_at], (long long)s->P);
    }
    for (int64_t g = 0; g < n_genes; g++) {
        const int64_t b = row_ptr[g], e = row_ptr[g + 1];
        for (int64_t j = b; j < e; j += item_pairs) {
            Item it;
            it.gene = (int32_t)g; it.begin = (int32_t)j;
            it.count = (int32_t)((e - j) < item_pairs ? (e - j) : item_pairs); it.pad = 0;
            items[n_items++] = it;
        }
    }
    EPX_CUDA(s->probe_idx.reserve((size_t)Np));
    EPX_CUDA(s->pair_slot.reserve((size_t)Np));
    EPX_CUDA(s->pair_gene.reserve((size_t)Np));
    EPX_CUDA(s->uniq_probe.reserve(n_uniq));
    EPX_CUDA(s->items.reserve(n_items));
    if (Np) {
        EPX_CUDA(cudaMemcpyAsync(s->probe_idx.p, h_pi, (size_t)Np * 4, cudaMemcpyHostToDevice, s->stream));
        EPX_CUDA(cudaMemcpyAsync(s->pair_slot.p, pair_slot, (size_t)Np * 4, cudaMemcpyHostToDevice, s->stream));
        EPX_CUDA(cudaMemcpyAsync(s->pair_gene.p, pair_gene, (size_t)Np * 4, cudaMemcpyHostToDevice, s->stream));
    }
    if (n_uniq) EPX_CUDA(cudaMemcpyAsync(s->uniq_probe.p, uniq, n_uniq * 4, cudaMemcpyHostToDevice, s->stream));
    if (n_items) EPX_CUDA(cudaMemcpyAsync(s->items.p, items, n_items * sizeof(Item), cudaMemcpyHostToDevice, s->stream));
    s->item_begin.resize(n_items + 1);
    for (size_t i = 0; i < n_items; i++) s->item_begin[i] = items[i].begin;
    s->item_begin[n_items] = (int32_t)Np;
    s->h_row_ptr.assign(row_ptr, row_ptr + n_genes + 1);
    s->h_probe_idx = h_pi;          /* host copies for the group-list checks: they live in the staging block */
    s->h_pair_gene = pair_gene;
    s->index_P = s->P;
    /* the copies read the staging block: it is rewritten by the next set_index only, which runs after this stream has
     * been synchronised by the run / fetch in between -- but a caller may set two indices back to back */
    EPX_CUDA(cudaStreamSynchronize(s->stream));
    s->n_groups = -1;
    s->prepared = false;
    s->Np = Np; s->U = (int64_t)n_uniq; s->n_items = (int64_t)n_items; s->G_index = n_genes;
    return EPX_OK;
}

int epx_session_set_index(epx_session *s, const int64_t *row_ptr, const int32_t *probe_idx, int64_t n_genes,
                          char *err, size_t errlen)
{
    return set_index_impl(s, row_ptr, probe_idx, n_genes, false, err, errlen);
}

/* Tables that depend on the sample count alone, built once per n on the host with the same epx_math.h code the
 * kernels use: the KS p-value lookup for two groups of equal size m (index d = max|cA - cB|, numerator d*m)
 * and the Pearson p-value table of epx_pearson_h. */
static int build_ks_lut(epx_session *s, int n, int m, cudaStream_t st, char *err, size_t errlen)
{
    if (s->lut_m == m && s->pt_n > 0 && s->lut_n == n) return EPX_OK;
    {
        const double df = (double)n - 2.0;
        const int N = epx_pearson_grid(df > 1.0 ? df : 1.0);
        std::vector<double> tab((size_t)N + 1);
        for (int i = 0; i <= N; i++) tab[(size_t)i] = df >= 1.0 ? epx_pearson_h((double)i / (double)N, df) : 0.0;
        EPX_CUDA(s->pt_tab.reserve(tab.size()));
        EPX_CUDA(cudaMemcpyAsync(s->pt_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st));
        EPX_CUDA(cudaStreamSynchronize(st));
        s->pt_n = N;
        s->lut_n = n;
    }
    std::vector<double> lut((size_t)2 * (m + 1)), u((size_t)m + 2);
    const bool exact_ok = (double)m * (double)m < 10000.0;
    for (int d = 0; d <= m; d++) {
        const double dnum = (double)d * (double)m;
        lut[(size_t)d] = exact_ok ? epx_ks_p_exact(dnum, m, m, u.data()) : EPX_NAN;
        lut[(size_t)(m + 1 + d)] = epx_ks_p_asym(dnum, (double)m, (double)m);
    }
    EPX_CUDA(s->lut.reserve(lut.size()));
    EPX_CUDA(cudaMemcpyAsync(s->lut.p, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice, st));
    EPX_CUDA(cudaStreamSynchronize(st));
    s->lut_m = m;
    return EPX_OK;
}

/* The Welch p-value table of the pair kernel's t mode (epx_welch_p_tab): h(r; df) on a (df, r) grid covering
 * m - 1 <= df <= 2(m - 1), built with the same epx_math.h code the kernels use, once per group size. */
static int build_welch_table(epx_session *s, int m, cudaStream_t st, char *err, size_t errlen)
{
    if (m < 30) { s->wt_m = -1; return EPX_OK; }      /* small groups keep the continued fraction */
    if (s->wt_m == m) return EPX_OK;
    int N, K;
    double df0, ddf;
    epx_welch_grid(m, &N, &K, &df0, &ddf);
    const int ld = (N + 4) & ~3;
    std::vector<double> tab((size_t)K * (size_t)ld, 0.0);
    std::vector<std::thread> th;
    const int nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    for (int w = 0; w < nt; w++)
        th.emplace_back([&, w]() {
            for (int k = w; k < K; k += nt)
                for (int i = 0; i <= N; i++) tab[(size_t)k * ld + i] = epx_pearson_h((double)i / (double)N, df0 + k * ddf);
        });
    for (auto &t : th) t.join();
    EPX_CUDA(s->wt_tab.reserve(tab.size()));
    EPX_CUDA(cudaMemcpyAsync(s->wt_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st));
    EPX_CUDA(cudaStreamSynchronize(st));
    s->wt_m = m; s->wt_N = N; s->wt_ld = ld; s->wt_K = K; s->wt_df0 = df0; s->wt_ddf = ddf;
    return EPX_OK;
}

int epx_session_run_prepare(epx_session *s, const epx_options *opt_in, void *stream, char *err, size_t errlen)
{
    if (!s) return fail(err, errlen, EPX_EINVAL, "session is NULL");
    epx_options opt;
    epx_options_default(&opt);
    if (opt_in) {
        size_t sz = opt_in->struct_size > 0 && (size_t)opt_in->struct_size < sizeof(opt) ? (size_t)opt_in->struct_size : sizeof(opt);
        memcpy(&opt, opt_in, sz);
    }
    if (!s->meth || !s->expr.p) return fail(err, errlen, EPX_ESTATE, "methylation and expression must be set before run");
    if (s->G_index != s->G) return fail(err, errlen, EPX_ESTATE, "index covers %lld genes, expression has %lld", (long long)s->G_index, (long long)s->G);
    if (s->n_meth != s->n) return fail(err, errlen, EPX_EINVAL, "sample counts differ: expression %lld, methylation %lld", (long long)s->n, (long long)s->n_meth);
    const int n = (int)s->n;
    if (opt.strict_n3 && n % 3 != 0)
        return fail(err, errlen, EPX_ESTRICT, "arguments imply differing number of rows: %d, %d", n, 3 * (n / 3));
    EPX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    const int m = n / 3;
    const int64_t G = s->G, Np = s->Np;

    s->ldl = round_up(n, 16); s->ldx = round_up(n, 2);
    s->ldo = n <= MAX_WARP_SAMPLES ? 32 * next_pow2((n + 31) / 32) : round_up(n, 16); /* every lane of the walk reads real or pad entries */
    EPX_CUDA(s->perm.reserve((size_t)(G * n)));
    EPX_CUDA(s->lab8.reserve((size_t)(G * s->ldl)));
    EPX_CUDA(s->xc.reserve((size_t)(G * s->ldx)));
    EPX_CUDA(s->gene_aux.reserve((size_t)(G * 4)));
    EPX_CUDA(s->nuniq.reserve((size_t)G));
    EPX_CUDA(s->ord.reserve((size_t)(s->U * s->ldo)));
    EPX_CUDA(s->gene.reserve((size_t)(G * GENE_COLS)));
    EPX_CUDA(s->gene_na.reserve((size_t)G));
    EPX_CUDA(s->gene_dnum.reserve((size_t)(G * 3)));
    EPX_CUDA(s->pair.reserve((size_t)(Np * PAIR_COLS)));
    EPX_CUDA(s->pair_na.reserve((size_t)Np));
    EPX_CUDA(s->pair_stat.reserve((size_t)(Np * 3)));
    EPX_CUDA(s->pair_dnum.reserve((size_t)(Np * 3)));
    int rc = build_ks_lut(s, n, m, st, err, errlen);
    if (rc) return rc;
    if (opt.test_meth_t) {
        rc = build_welch_table(s, m, st, err, errlen);
        if (rc) return rc;
    }

    GeneParams gp;
    gp.expr = s->expr.p; gp.n = n; gp.G = (int)G; gp.npad = next_pow2(n); gp.m = m;
    gp.perm = s->perm.p; gp.lab8 = s->lab8.p; gp.ldl = s->ldl; gp.xc = s->xc.p; gp.ldx = s->ldx;
    gp.gene_aux = s->gene_aux.p; gp.nuniq = s->nuniq.p; gp.one = 1u;

    GeneTestParams tp;
    tp.expr = s->expr.p; tp.perm = s->perm.p; tp.n = n; tp.G = (int)G; tp.nuniq = s->nuniq.p;
    tp.counts = s->scal.p; tp.ref_gene = s->scal.p + 3; tp.status = s->scal.p + 4;
    tp.test_t = opt.test_expr_t; tp.data_linear = opt.data_linear; tp.fc_from_means = opt.fc_from_means;
    tp.gene = s->gene.p; tp.gene_na = s->gene_na.p; tp.gene_dnum = s->gene_dnum.p;

    RankParams rp;
    rp.meth = s->meth; rp.ld = s->ld; rp.uniq_probe = s->uniq_probe.p; rp.U = s->U; rp.n = n; rp.npad = next_pow2(n);
    rp.ord = s->ord.p; rp.ldo = s->ldo; rp.one = 1u; rp.queue = s->scal.p + 10; rp.gene_queue = s->scal.p + 8;
    /* n <= 512: the expression rows go through the row kernel too (launch_gene_rows), whose last warp picks the
     * bin.var reference gene */
    const bool fuse_sort = probe_rank_takes_genes(n);
    rp.genes = gp;
    rp.nuniq_in = s->nuniq.p; rp.counts = tp.counts; rp.ref_gene = tp.ref_gene; rp.status = tp.status;
    rp.best = (unsigned long long *)(s->scal.p + 12);
    rp.full_rows = s->scal.p + 14;

    PairParams pp;
    pp.meth = s->meth; pp.ld = s->ld; pp.ord = s->ord.p; pp.ldo = s->ldo; pp.probe_idx = s->probe_idx.p;
    pp.pair_slot = s->pair_slot.p; pp.xc = s->xc.p; pp.ldx = s->ldx; pp.lab8 = s->lab8.p; pp.ldl = s->ldl;
    pp.perm = s->perm.p; pp.gene_aux = s->gene_aux.p; pp.items = s->items.p; pp.n_items = (int)s->n_items; pp.queue = s->scal.p + 6; pp.pair_gene = s->pair_gene.p; pp.n_pairs = (int)Np; pp.pair_begin = 0;
    pp.n = n; pp.m = m; pp.test_t = opt.test_meth_t;
    pp.exact_ok = ((double)m * (double)m < 10000.0) ? 1 : 0;
    pp.warp_smem = 0;
    pp.pt_tab = s->pt_tab.p; pp.pt_n = s->pt_n;
    pp.lut_exact = s->lut.p; pp.lut_asym = s->lut.p + (m + 1);
    const bool wt = opt.test_meth_t && s->wt_m == m;
    pp.wt_tab = wt ? s->wt_tab.p : nullptr;
    pp.wt_N = s->wt_N; pp.wt_ld = s->wt_ld; pp.wt_K = s->wt_K; pp.wt_df0 = s->wt_df0; pp.wt_inv_ddf = 1.0 / s->wt_ddf;
    pp.pair = s->pair.p; pp.pair_na = s->pair_na.p; pp.pair_stat = s->pair_stat.p; pp.pair_dnum = s->pair_dnum.p;

    /* The gene-level kernels depend on the expression matrix only and the probe ranking on the methylation
     * matrix only: they run concurrently (side stream / main stream) and join before the pair kernel.
     * options.reserved[0] & 1 serialises them (clean per-kernel timings for profiling). */
    const bool overlap = !(opt.reserved[0] & 1);
    cudaStream_t gs = overlap ? s->side : st;
    s->launches = 0;
    EPX_CUDA(cudaEventRecord(s->ev[0], st));
    if (overlap) EPX_CUDA(cudaStreamWaitEvent(gs, s->ev[0], 0));
    EPX_CUDA(cudaEventRecord(s->ev[5], gs));
    if (fuse_sort) { EPX_CUDA(launch_gene_rows(rp, s->sm_count, gs)); s->launches++; } /* order + reference pick */
    else { EPX_CUDA(launch_gene_sort(gp, s->sm_count, gs)); s->launches++; }
    EPX_CUDA(cudaEventRecord(s->ev[1], gs));
    if (!fuse_sort) { EPX_CUDA(launch_pick_reference(tp, gs)); s->launches++; }
    EPX_CUDA(launch_gene_tests(tp, s->sm_count, gs)); s->launches++;
    EPX_CUDA(cudaEventRecord(s->ev[2], gs));
    EPX_CUDA(cudaEventRecord(s->ev[6], st));
    if (s->U > 0) { EPX_CUDA(launch_probe_rank(rp, s->sm_count, st)); s->launches++; }
    EPX_CUDA(cudaEventRecord(s->ev[3], st));
    /* the pair kernel needs the gene order/labels/centred expression (gene sort) but not the gene-level tests: it
     * waits for ev[1] only, and the main stream joins the rest of the side stream after the pair kernel (run_pairs) */
    if (overlap) EPX_CUDA(cudaStreamWaitEvent(st, s->ev[1], 0));
    s->join_side = overlap;
    s->pp = pp;
    s->prepared = true;
    s->ran = false;
    return EPX_OK;
}

static void chunk_items(const epx_session *s, int chunk, int n_chunks, int64_t *i0, int64_t *i1)
{
    *i0 = s->n_items * chunk / n_chunks;
    *i1 = s->n_items * (chunk + 1) / n_chunks;
}

int epx_session_pairs_chunk(epx_session *s, int chunk, int n_chunks, int64_t *first_pair, int64_t *n_pairs, char *err,
                            size_t errlen)
{
    if (!s || n_chunks < 1 || chunk < 0 || chunk >= n_chunks) return fail(err, errlen, EPX_EINVAL, "bad chunk %d of %d", chunk, n_chunks);
    if (s->G_index < 0) return fail(err, errlen, EPX_ESTATE, "no index");
    int64_t i0, i1;
    chunk_items(s, chunk, n_chunks, &i0, &i1);
    const int64_t b = s->n_items ? s->item_begin[(size_t)i0] : 0, e = s->n_items ? s->item_begin[(size_t)i1] : 0;
    if (first_pair) *first_pair = b;
    if (n_pairs) *n_pairs = e - b;
    return EPX_OK;
}

int epx_session_run_pairs(epx_session *s, int chunk, int n_chunks, void *stream, char *err, size_t errlen)
{
    if (!s || n_chunks < 1 || chunk < 0 || chunk >= n_chunks) return fail(err, errlen, EPX_EINVAL, "bad chunk %d of %d", chunk, n_chunks);
    if (!s->prepared) return fail(err, errlen, EPX_ESTATE, "run_pairs before run_prepare");
    EPX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    int64_t i0, i1;
    chunk_items(s, chunk, n_chunks, &i0, &i1);
    if (chunk == 0) EPX_CUDA(cudaEventRecord(s->ev[7], st));
    if (i1 > i0) {
        PairParams pp = s->pp;
        pp.items = s->pp.items + i0;
        pp.n_items = (int)(i1 - i0);
        pp.pair_begin = s->item_begin[(size_t)i0];
        pp.n_pairs = s->item_begin[(size_t)i1];
        EPX_CUDA(launch_pair_stats(pp, s->sm_count, st));
        s->launches++;
    }
    if (chunk == n_chunks - 1) {
        if (s->join_side) EPX_CUDA(cudaStreamWaitEvent(st, s->ev[2], 0)); /* gene tests (side stream) are part of the run */
        EPX_CUDA(cudaEventRecord(s->ev[4], st));
        s->ran = true;
    }
    return EPX_OK;
}

int epx_session_run(epx_session *s, const epx_options *opt, void *stream, char *err, size_t errlen)
{
    int rc = epx_session_run_prepare(s, opt, stream, err, errlen);
    if (rc) return rc;
    return epx_session_run_pairs(s, 0, 1, stream, err, errlen);
}

int epx_session_wait(epx_session *s, void *stream, char *err, size_t errlen)
{
    if (!s) return fail(err, errlen, EPX_EINVAL, "session is NULL");
    EPX_CUDA(cudaSetDevice(s->device));
    EPX_CUDA(cudaStreamSynchronize(stream ? (cudaStream_t)stream : s->stream));
    return EPX_OK;
}

int epx_session_fetch(epx_session *s, epx_result *out, void *stream, char *err, size_t errlen)
{
    if (!s || !out) return fail(err, errlen, EPX_EINVAL, "NULL argument");
    if (!s->ran) return fail(err, errlen, EPX_ESTATE, "fetch before run");
    EPX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    int32_t scal[5];
    EPX_CUDA(cudaMemcpyAsync(scal, s->scal.p, sizeof(scal), cudaMemcpyDeviceToHost, st));
    const size_t Np = (size_t)s->Np, G = (size_t)s->G, n = (size_t)s->n;
    if (out->pair && Np) EPX_CUDA(cudaMemcpyAsync(out->pair, s->pair.p, Np * PAIR_COLS * 8, cudaMemcpyDeviceToHost, st));
    if (out->pair_na && Np) EPX_CUDA(cudaMemcpyAsync(out->pair_na, s->pair_na.p, Np * 2, cudaMemcpyDeviceToHost, st));
    if (out->pair_stat && Np) EPX_CUDA(cudaMemcpyAsync(out->pair_stat, s->pair_stat.p, Np * 3 * 8, cudaMemcpyDeviceToHost, st));
    if (out->pair_ks_dnum && Np) EPX_CUDA(cudaMemcpyAsync(out->pair_ks_dnum, s->pair_dnum.p, Np * 3 * 4, cudaMemcpyDeviceToHost, st));
    if (out->gene) EPX_CUDA(cudaMemcpyAsync(out->gene, s->gene.p, G * GENE_COLS * 8, cudaMemcpyDeviceToHost, st));
    if (out->gene_na) EPX_CUDA(cudaMemcpyAsync(out->gene_na, s->gene_na.p, G * 2, cudaMemcpyDeviceToHost, st));
    if (out->gene_ks_dnum) EPX_CUDA(cudaMemcpyAsync(out->gene_ks_dnum, s->gene_dnum.p, G * 3 * 4, cudaMemcpyDeviceToHost, st));
    if (out->perm) EPX_CUDA(cudaMemcpyAsync(out->perm, s->perm.p, G * n * 2, cudaMemcpyDeviceToHost, st));
    EPX_CUDA(cudaStreamSynchronize(st));
    out->counts[0] = scal[0]; out->counts[1] = scal[1]; out->counts[2] = scal[2];
    out->ref_gene = scal[3];
    if (scal[4] == ST_BREAKS_NOT_UNIQUE)
        return fail(err, errlen, EPX_EBREAKS, "bin.var on reference gene %d: 'breaks' are not unique", scal[3]);
    return EPX_OK;
}

int epx_session_timing(epx_session *s, epx_timing *out, char *err, size_t errlen)
{
    if (!s || !out) return fail(err, errlen, EPX_EINVAL, "NULL argument");
    if (!s->ran) return fail(err, errlen, EPX_ESTATE, "timing before run");
    EPX_CUDA(cudaSetDevice(s->device));
    EPX_CUDA(cudaEventSynchronize(s->ev[4]));
    memset(out, 0, sizeof(*out));
    EPX_CUDA(cudaEventElapsedTime(&out->total_ms, s->ev[0], s->ev[4]));
    EPX_CUDA(cudaEventElapsedTime(&out->gene_sort_ms, s->ev[5], s->ev[1]));
    EPX_CUDA(cudaEventElapsedTime(&out->gene_tests_ms, s->ev[1], s->ev[2]));
    EPX_CUDA(cudaEventElapsedTime(&out->probe_rank_ms, s->ev[6], s->ev[3]));
    EPX_CUDA(cudaEventElapsedTime(&out->pair_stats_ms, s->ev[7], s->ev[4]));
    out->n_pairs = s->Np; out->n_unique_probes = s->U; out->n_genes = s->G; out->n_samples = s->n;
    out->algorithmic_bytes = s->Np * (8 * s->n + 88) + s->G * (10 * s->n + 48);
    out->kernel_launches = s->launches;
    EPX_CUDA(cudaMemcpy(&out->sort_full_rows, s->scal.p + 14, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return EPX_OK;
}

int epx_session_result_device_ptr(epx_session *s, int which, void **ptr, int64_t *bytes, char *err, size_t errlen)
{
    if (!s || !ptr) return fail(err, errlen, EPX_EINVAL, "NULL argument");
    if (!s->ran && !s->prepared) return fail(err, errlen, EPX_ESTATE, "no results yet");
    const int64_t Np = s->Np, G = s->G, n = s->n;
    void *p = nullptr; int64_t b = 0;
    switch (which) {
    case 0: p = s->pair.p; b = Np * PAIR_COLS * 8; break;
    case 1: p = s->pair_na.p; b = Np * 2; break;
    case 2: p = s->pair_stat.p; b = Np * 24; break;
    case 3: p = s->pair_dnum.p; b = Np * 12; break;
    case 4: p = s->gene.p; b = G * GENE_COLS * 8; break;
    case 5: p = s->gene_na.p; b = G * 2; break;
    case 6: p = s->gene_dnum.p; b = G * 12; break;
    case 7: p = s->perm.p; b = G * n * 2; break;
    default: return fail(err, errlen, EPX_EINVAL, "unknown table %d", which);
    }
    *ptr = p;
    if (bytes) *bytes = b;
    return EPX_OK;
}

/* ---------------------------------------------------------------- pooled probe groups */

int epx_session_run_groups(epx_session *s, const int64_t *group_ptr, const int32_t *group_pairs, int64_t n_groups,
                           void *stream, char *err, size_t errlen)
{
    if (!s || !group_ptr || n_groups < 0) return fail(err, errlen, EPX_EINVAL, "bad group arguments");
    if (!s->prepared && !s->ran) return fail(err, errlen, EPX_ESTATE, "run_groups before run / run_prepare");
    if (n_groups > 0x7fffffff || group_ptr[0] != 0) return fail(err, errlen, EPX_EINVAL, "bad group_ptr");
    const int64_t L = group_ptr[n_groups];
    if (L < 0 || L > 0x7fffffff || (L > 0 && !group_pairs)) return fail(err, errlen, EPX_EINVAL, "bad group pair list");
    EPX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    const int n = (int)s->n, m = n / 3;
    const int cap = group_block_cap(n), rpb = cap / n > 0 ? cap / n : 1;

    std::vector<GroupDesc> desc((size_t)n_groups);
    std::vector<BlockDesc> blocks;
    int64_t elem = 0;
    int max_blocks = 1;
    for (int64_t q = 0; q < n_groups; q++) {
        const int64_t b = group_ptr[q], e = group_ptr[q + 1];
        if (e < b || e > L) return fail(err, errlen, EPX_EINVAL, "group_ptr is not monotone at group %lld", (long long)q);
        GroupDesc &d = desc[(size_t)q];
        d.k = (int32_t)(e - b); d.first = (int32_t)b; d.elem = elem; d.gene = 0; d.nblocks = 0;
        if (d.k == 0) continue;
        for (int64_t j = b; j < e; j++) {
            const int32_t pr = group_pairs[j];
            if (pr < 0 || pr >= s->Np) return fail(err, errlen, EPX_EINVAL, "group %lld names pair %d outside [0,%lld)", (long long)q, pr, (long long)s->Np);
            if (s->h_probe_idx[(size_t)pr] < 0) return fail(err, errlen, EPX_EINVAL, "group %lld names pair %d, which has no data row", (long long)q, pr);
            if (s->h_pair_gene[(size_t)pr] != s->h_pair_gene[(size_t)group_pairs[b]])
                return fail(err, errlen, EPX_EINVAL, "group %lld mixes genes %d and %d", (long long)q, s->h_pair_gene[(size_t)group_pairs[b]], s->h_pair_gene[(size_t)pr]);
        }
        d.gene = s->h_pair_gene[(size_t)group_pairs[b]];
        d.nblocks = (d.k + rpb - 1) / rpb;
        if (d.nblocks > max_blocks) max_blocks = d.nblocks;
        for (int r0 = 0; r0 < d.k; r0 += rpb) {
            BlockDesc bd;
            bd.group = (int32_t)q; bd.r0 = r0; bd.runs = d.k - r0 < rpb ? d.k - r0 : rpb; bd.pad = 0;
            blocks.push_back(bd);
        }
        elem += (int64_t)d.k * n;
    }
    /* merge tiles, groups with the most runs first: pass p merges the groups with more than 2^p runs = a prefix */
    std::vector<int32_t> multi;
    for (int64_t q = 0; q < n_groups; q++) if (desc[(size_t)q].nblocks > 1) multi.push_back((int32_t)q);
    std::sort(multi.begin(), multi.end(), [&](int32_t a, int32_t b) {
        return desc[(size_t)a].nblocks != desc[(size_t)b].nblocks ? desc[(size_t)a].nblocks > desc[(size_t)b].nblocks : a < b; });
    const int64_t tile = 256 * 16; /* MERGE_TILE */
    std::vector<TileDesc> tiles;
    int passes = 0;
    while ((1 << passes) < max_blocks) passes++;
    std::vector<int64_t> pass_tiles((size_t)passes, 0);
    for (int32_t q : multi) {
        const GroupDesc &d = desc[(size_t)q];
        const int64_t N = (int64_t)d.k * n, nt = (N + tile - 1) / tile;
        for (int64_t t = 0; t < nt; t++) { TileDesc td; td.group = q; td.tile = (int32_t)t; tiles.push_back(td); }
        for (int p = 0; p < passes; p++) if (d.nblocks > (1 << p)) pass_tiles[(size_t)p] = (int64_t)tiles.size();
    }
    /* groups that do not finish inside k_group_blocksort: several blocks (moments + merges + walk), or empty (NA row) */
    std::vector<int32_t> rest;
    for (int64_t q = 0; q < n_groups; q++) if (desc[(size_t)q].nblocks != 1) rest.push_back((int32_t)q);
    if (tiles.size() > 0x7fffffffu || blocks.size() > 0x7fffffffu) return fail(err, errlen, EPX_EINVAL, "too many group blocks");

    EPX_CUDA(s->g_desc.reserve(desc.size()));
    EPX_CUDA(s->g_blocks.reserve(blocks.size()));
    EPX_CUDA(s->g_tiles.reserve(tiles.size()));
    EPX_CUDA(s->g_pairs.reserve((size_t)L));
    EPX_CUDA(s->g_multi.reserve(rest.size()));
    EPX_CUDA(s->g_guard.reserve((size_t)n_groups * 3));
    EPX_CUDA(s->grp.reserve((size_t)n_groups * PAIR_COLS));
    EPX_CUDA(s->grp_stat.reserve((size_t)n_groups * 3));
    EPX_CUDA(s->grp_na.reserve((size_t)n_groups));
    EPX_CUDA(s->grp_dnum.reserve((size_t)n_groups * 3));
    for (int i = 0; i < 2; i++) {
        if (i == 1 && passes == 0) break;
        EPX_CUDA(s->g_val[i].reserve((size_t)elem));
        EPX_CUDA(s->g_lab[i].reserve((size_t)elem));
    }
    if (!desc.empty()) EPX_CUDA(cudaMemcpyAsync(s->g_desc.p, desc.data(), desc.size() * sizeof(GroupDesc), cudaMemcpyHostToDevice, st));
    if (!blocks.empty()) EPX_CUDA(cudaMemcpyAsync(s->g_blocks.p, blocks.data(), blocks.size() * sizeof(BlockDesc), cudaMemcpyHostToDevice, st));
    if (!tiles.empty()) EPX_CUDA(cudaMemcpyAsync(s->g_tiles.p, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, st));
    if (!rest.empty()) EPX_CUDA(cudaMemcpyAsync(s->g_multi.p, rest.data(), rest.size() * 4, cudaMemcpyHostToDevice, st));
    if (L) EPX_CUDA(cudaMemcpyAsync(s->g_pairs.p, group_pairs, (size_t)L * 4, cudaMemcpyHostToDevice, st));
    EPX_CUDA(cudaStreamSynchronize(st)); /* the vectors die with this scope */

    GroupParams gp;
    gp.meth = s->meth; gp.ld = s->ld; gp.probe_idx = s->probe_idx.p; gp.group_pairs = s->g_pairs.p;
    gp.pair_slot = s->pair_slot.p; gp.ord = s->ord.p; gp.ldo = s->ldo;
    gp.groups = s->g_desc.p; gp.blocks = s->g_blocks.p; gp.tiles = s->g_tiles.p;
    gp.multi = s->g_multi.p; gp.n_multi = (int)rest.size();
    gp.n_groups = (int)n_groups; gp.n_blocks = (int)blocks.size();
    gp.n = n; gp.m = m; gp.cap = cap; gp.test_t = s->pp.test_t;
    gp.perm = s->perm.p; gp.lab8 = s->lab8.p; gp.ldl = s->ldl; gp.xc = s->xc.p; gp.ldx = s->ldx; gp.gene_aux = s->gene_aux.p;
    gp.val[0] = s->g_val[0].p; gp.val[1] = s->g_val[1].p; gp.lab[0] = s->g_lab[0].p; gp.lab[1] = s->g_lab[1].p;
    gp.guard = s->g_guard.p; gp.grp = s->grp.p; gp.grp_na = s->grp_na.p; gp.grp_stat = s->grp_stat.p; gp.grp_dnum = s->grp_dnum.p;
    gp.one = 1u;
    EPX_CUDA(launch_group_moments(gp, st)); s->launches++;
    EPX_CUDA(launch_group_blocksort(gp, st)); s->launches++;
    for (int p = 0; p < passes; p++) {
        EPX_CUDA(launch_group_merge(gp, p, ((int64_t)rpb * n) << p, (int)pass_tiles[(size_t)p], st));
        s->launches++;
    }
    EPX_CUDA(launch_group_walk(gp, st)); s->launches++;
    s->n_groups = n_groups;
    return EPX_OK;
}

int epx_session_fetch_groups(epx_session *s, epx_group_result *out, void *stream, char *err, size_t errlen)
{
    if (!s || !out) return fail(err, errlen, EPX_EINVAL, "NULL argument");
    if (s->n_groups < 0) return fail(err, errlen, EPX_ESTATE, "fetch_groups before run_groups");
    EPX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    const size_t Q = (size_t)s->n_groups;
    if (out->group && Q) EPX_CUDA(cudaMemcpyAsync(out->group, s->grp.p, Q * PAIR_COLS * 8, cudaMemcpyDeviceToHost, st));
    if (out->group_na && Q) EPX_CUDA(cudaMemcpyAsync(out->group_na, s->grp_na.p, Q * 2, cudaMemcpyDeviceToHost, st));
    if (out->group_stat && Q) EPX_CUDA(cudaMemcpyAsync(out->group_stat, s->grp_stat.p, Q * 24, cudaMemcpyDeviceToHost, st));
    if (out->group_ks_dnum && Q) EPX_CUDA(cudaMemcpyAsync(out->group_ks_dnum, s->grp_dnum.p, Q * 24, cudaMemcpyDeviceToHost, st));
    EPX_CUDA(cudaStreamSynchronize(st));
    return EPX_OK;
}

int epx_analyze(epx_session *s, const double *expr, int64_t n_genes, int64_t n_samples, const int64_t *row_ptr,
                const int32_t *probe_idx, const epx_options *opt, epx_result *out, char *err, size_t errlen)
{
    /* the expression matrix is on its way to the device while the host builds the index */
    int rc = set_expression_begin(s, expr, n_genes, n_samples, 0, err, errlen);
    if (rc) return rc;
    rc = epx_session_set_index(s, row_ptr, probe_idx, n_genes, err, errlen);
    if (rc) { cudaStreamSynchronize(s->stream); return rc; }
    rc = set_expression_finish(s, n_genes, n_samples, err, errlen);
    if (rc) return rc;
    rc = epx_session_run(s, opt, nullptr, err, errlen);
    if (rc) return rc;
    return epx_session_fetch(s, out, nullptr, err, errlen);
}

/* The caller's matrix as the device sees it, when the device can read it in place: page-locked host memory
 * (epx_host_alloc, cudaHostAlloc, cudaHostRegister) -- or device memory, which the same kernel reads just as well.
 * NULL for pageable memory. Both ends of the matrix must lie in the same mapping. */
static const double *rows_readable_by_device(const double *rows, int64_t n_probes, int64_t n_samples, int64_t ld)
{
    const double *last = rows + (n_probes - 1) * ld + (n_samples - 1);
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, rows) != cudaSuccess || cudaPointerGetAttributes(&a1, last) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if ((a0.type != cudaMemoryTypeHost && a0.type != cudaMemoryTypeDevice) || a1.type != a0.type) return nullptr;
    if (!a0.devicePointer || !a1.devicePointer) return nullptr;
    if ((const char *)a1.devicePointer - (const char *)a0.devicePointer != (const char *)last - (const char *)rows) return nullptr;
    return (const double *)a0.devicePointer;
}

/* epx_analyze_host / epx_multi_analyze_host, first half: the session gets a device matrix of the caller's shape with no
 * valid row in it yet (the index is built against its row count before the rows it names are fetched) */
static int host_rows_begin(epx_session *s, int64_t n_probes, int64_t n_samples, char *err, size_t errlen)
{
    const int64_t dld = round_up(n_samples, 4);
    /* no usable matrix until the rows have arrived and passed the finite check */
    s->meth = nullptr; s->prepared = false; s->ran = false; s->n_groups = -1;
    EPX_CUDA(s->meth_own.reserve((size_t)(n_probes * dld)));
    if (s->index_P != n_probes) s->G_index = -1; /* as in set_methylation_*: the index depends on the row count only */
    s->P = n_probes; s->n_meth = n_samples; s->ld = dld; s->meth_partial = true;
    return EPX_OK;
}

/* second half, after the index: the rows it names (s->uniq_probe, ascending), straight from the caller's memory into
 * their places, checked for NaN / infinity; blocking */
static int host_rows_fetch(epx_session *s, const double *src, int64_t ld, char *err, size_t errlen)
{
    EPX_CUDA(arm_finite_check(s, s->stream));
    EPX_CUDA(launch_gather_host_rows(src, ld, s->uniq_probe.p, s->U, s->n_meth, s->meth_own.p, s->ld, s->first_bad.p, s->stream));
    EPX_CUDA(cudaStreamSynchronize(s->stream));
    int rc = finite_check_result(s, "methylation", "probe", s->n_meth, err, errlen);
    if (rc) return rc;
    s->meth = s->meth_own.p;
    return EPX_OK;
}

int epx_analyze_host(epx_session *s, const double *meth_rows, int64_t n_probes, int64_t ld, const double *expr,
                     int64_t n_genes, int64_t n_samples, const int64_t *row_ptr, const int32_t *probe_idx,
                     const epx_options *opt, epx_result *out, char *err, size_t errlen)
{
    if (!s || !meth_rows || n_probes < 1 || ld < n_samples) return fail(err, errlen, EPX_EINVAL, "bad methylation arguments");
    int rc = check_samples(s, n_samples, err, errlen);
    if (rc) return rc;
    EPX_CUDA(cudaSetDevice(s->device));
    const double *src = getenv("EPX_NO_ZEROCOPY") ? nullptr : rows_readable_by_device(meth_rows, n_probes, n_samples, ld);
    if (!src) {
        /* pageable memory: the device cannot fetch rows by itself; whole matrix through the staged upload */
        rc = epx_session_set_methylation_rows(s, meth_rows, n_probes, n_samples, ld, err, errlen);
        if (rc) return rc;
        return epx_analyze(s, expr, n_genes, n_samples, row_ptr, probe_idx, opt, out, err, errlen);
    }
    const bool trace = getenv("EPX_TRACE") != nullptr; /* host-side phase times, to stderr */
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    rc = host_rows_begin(s, n_probes, n_samples, err, errlen);
    if (rc) return rc;
    /* the expression matrix is on its way to the device while the host builds the index */
    rc = set_expression_begin(s, expr, n_genes, n_samples, 0, err, errlen);
    if (rc) return rc;
    rc = set_index_impl(s, row_ptr, probe_idx, n_genes, true, err, errlen);
    if (rc) { cudaStreamSynchronize(s->stream); return rc; }
    rc = set_expression_finish(s, n_genes, n_samples, err, errlen);
    if (rc) return rc;
    const double t_index = ms_since(t_start);
    const auto t_rows = std::chrono::steady_clock::now();
    rc = host_rows_fetch(s, src, ld, err, errlen);
    if (rc) return rc;
    const double t_fetch = ms_since(t_rows);
    const auto t_run = std::chrono::steady_clock::now();
    rc = epx_session_run(s, opt, nullptr, err, errlen);
    if (rc) return rc;
    rc = epx_session_fetch(s, out, nullptr, err, errlen);
    if (trace)
        fprintf(stderr, "epx_analyze_host: expression upload + index %.2f ms, %lld rows of %lld fetched in %.2f ms (%.1f GB/s), "
                        "kernels + tables back %.2f ms\n", t_index, (long long)s->U, (long long)n_probes, t_fetch,
                (double)s->U * (double)n_samples * 8.0 / (t_fetch * 1e6), ms_since(t_run));
    return rc;
}


/* ---------------------------------------------------------------- several GPUs, one process */


int epx_multi_create(const int *devices, int n_devices, epx_multi **out, char *err, size_t errlen)
{
    if (!out || !devices || n_devices < 1) return fail(err, errlen, EPX_EINVAL, "bad device list");
    *out = nullptr;
    epx_multi *m = new (std::nothrow) epx_multi();
    if (!m) return fail(err, errlen, EPX_ENOMEM, "out of host memory");
    for (int i = 0; i < n_devices; i++) {
        epx_session *s = nullptr;
        int rc = epx_session_create(devices[i], &s, err, errlen);
        if (rc) { epx_multi_destroy(m); return rc; }
        m->sess.push_back(s);
    }
    /* one host thread per device drives its session at the same time: the index builds share the cores (their workers
     * meet at spin barriers, so more workers than cores costs scheduling quanta: 2 devices x 8 workers on 16 cores made
     * every call 3 ms longer) */
    int per = (int)std::thread::hardware_concurrency() / n_devices - 1;
    per = per < 1 ? 1 : (per > 8 ? 8 : per);
    for (epx_session *s : m->sess) s->host_threads = per;
    /* the expression matrix crosses PCIe once (to the first device); the others fetch it from there over NVLink */
    if (n_devices > 1) {
        for (int i = 0; i < n_devices; i++) {
            if (cudaSetDevice(devices[i]) != cudaSuccess) continue;
            for (int j = 0; j < n_devices; j++)
                if (j != i && cudaDeviceEnablePeerAccess(devices[j], 0) != cudaSuccess) cudaGetLastError(); /* already on, or no path: the copy is staged */
        }
        if (cudaSetDevice(devices[0]) != cudaSuccess || cudaEventCreateWithFlags(&m->ev_expr, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            m->ev_expr = nullptr;
        }
    }
    *out = m;
    return EPX_OK;
}

void epx_multi_destroy(epx_multi *m)
{
    if (!m) return;
    if (m->ev_expr && !m->sess.empty()) { cudaSetDevice(m->sess[0]->device); cudaEventDestroy(m->ev_expr); }
    for (epx_session *s : m->sess) epx_session_destroy(s);
    delete m;
}

int epx_multi_set_methylation_rows(epx_multi *m, const double *rows, int64_t n_probes, int64_t n_samples, int64_t ld,
                                   char *err, size_t errlen)
{
    if (!m) return fail(err, errlen, EPX_EINVAL, "NULL handle");
    m->sliced = false;
    m->rows_of.clear();
    return for_each_device(m, [&](int i, char *e, size_t el) {
        return epx_session_set_methylation_rows(m->sess[(size_t)i], rows, n_probes, n_samples, ld, e, el);
    }, err, errlen);
}

int epx_multi_set_methylation_cols(epx_multi *m, const double *const *cols, int64_t n_samples, int64_t n_probes,
                                   char *err, size_t errlen)
{
    if (!m) return fail(err, errlen, EPX_EINVAL, "NULL handle");
    m->sliced = false;
    m->rows_of.clear();
    return for_each_device(m, [&](int i, char *e, size_t el) {
        return epx_session_set_methylation_cols(m->sess[(size_t)i], cols, n_samples, n_probes, e, el);
    }, err, errlen);
}

int epx_multi_set_methylation_sliced(epx_multi *m, const double *rows, int64_t n_probes, int64_t n_samples, int64_t ld,
                                     const int64_t *row_ptr, const int32_t *probe_idx, int64_t n_genes,
                                     char *err, size_t errlen)
{
    if (!m) return fail(err, errlen, EPX_EINVAL, "NULL handle");
    if (!rows || !row_ptr || n_genes < 1 || n_probes < 1 || n_samples < 1 || ld < n_samples)
        return fail(err, errlen, EPX_EINVAL, "bad arguments");
    const int64_t Np = row_ptr[n_genes];
    if (Np > 0 && !probe_idx) return fail(err, errlen, EPX_EINVAL, "probe_idx is NULL");
    for (int64_t k = 0; k < Np; k++)
        if (probe_idx[k] >= n_probes) return fail(err, errlen, EPX_EINVAL, "probe_idx[%lld] = %d is outside the %lld methylation rows",
                                                  (long long)k, probe_idx[k], (long long)n_probes);
    const int W = (int)m->sess.size();
    const std::vector<int64_t> cut = gene_cuts(row_ptr, n_genes, W);
    m->sliced = false;
    m->rows_of.assign((size_t)W, std::vector<int32_t>());
    int rc = for_each_device(m, [&](int i, char *e, size_t el) -> int {
        const int64_t lo = row_ptr[cut[(size_t)i]], hi = row_ptr[cut[(size_t)i + 1]];
        std::vector<int32_t> &u = m->rows_of[(size_t)i];
        for (int64_t k = lo; k < hi; k++)
            if (probe_idx[k] >= 0) u.push_back(probe_idx[k]);
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        const size_t U = u.size() ? u.size() : 1;        /* a device without pairs still gets one (zero) row */
        double *compact = (double *)calloc(U * (size_t)n_samples, sizeof(double));
        if (!compact) return fail(e, el, EPX_ENOMEM, "out of host memory for %zu sliced rows", U);
        for (size_t j = 0; j < u.size(); j++)
            memcpy(compact + j * (size_t)n_samples, rows + (size_t)u[j] * (size_t)ld, (size_t)n_samples * sizeof(double));
        int r = epx_session_set_methylation_rows(m->sess[(size_t)i], compact, (int64_t)U, n_samples, n_samples, e, el);
        free(compact);
        return r;
    }, err, errlen);
    if (rc) { m->rows_of.clear(); return rc; }
    m->sliced = true;
    return EPX_OK;
}

/* epx_multi_analyze (meth_rows == NULL: every device holds its matrix already) and epx_multi_analyze_host (meth_rows =
 * the caller's page-locked matrix: every device fetches the rows ITS gene range names, after its index is built) */
static int multi_analyze_impl(epx_multi *m, const double *meth_rows, int64_t n_probes, int64_t ld, const double *expr,
                              int64_t n_genes, int64_t n_samples, const int64_t *row_ptr, const int32_t *probe_idx,
                              const epx_options *opt, epx_result *out, char *err, size_t errlen)
{
    if (!m || !expr || !row_ptr || !out || n_genes < 1) return fail(err, errlen, EPX_EINVAL, "bad arguments");
    const int W = (int)m->sess.size();
    const std::vector<int64_t> cut = gene_cuts(row_ptr, n_genes, W);
    m->pair_lo.clear();
    std::atomic<int> expr_ready{0}; /* 1: device 0 has queued its upload and recorded ev_expr; -1: it failed */
    int rc_all = for_each_device(m, [&](int i, char *e, size_t el) -> int {
        epx_session *s = m->sess[(size_t)i];
        if (meth_rows) {
            int rb = cudaSetDevice(s->device) == cudaSuccess ? host_rows_begin(s, n_probes, n_samples, e, el)
                                                             : fail(e, el, EPX_ECUDA, "cudaSetDevice(%d) failed", s->device);
            if (rb) {
                if (i == 0 && m->ev_expr) expr_ready.store(-1, std::memory_order_release); /* the other devices wait for this one's upload */
                return rb;
            }
        }
        const int64_t g0 = cut[(size_t)i], g1 = cut[(size_t)i + 1];
        const int64_t lo = row_ptr[g0], hi = row_ptr[g1];
        std::vector<int64_t> rp((size_t)n_genes + 1);
        for (int64_t g = 0; g <= n_genes; g++) {
            int64_t v = row_ptr[g];
            v = v < lo ? lo : (v > hi ? hi : v);
            rp[(size_t)g] = v - lo; /* all genes keep a row (gene kernels run everywhere); only [g0,g1) keep pairs */
        }
        epx_result r;
        memset(&r, 0, sizeof r);
        if (out->pair) r.pair = out->pair + lo * PAIR_COLS;
        if (out->pair_na) r.pair_na = out->pair_na + lo;
        if (out->pair_stat) r.pair_stat = out->pair_stat + lo * 3;
        if (out->pair_ks_dnum) r.pair_ks_dnum = out->pair_ks_dnum + lo * 3;
        if (i == 0) { r.gene = out->gene; r.gene_na = out->gene_na; r.gene_ks_dnum = out->gene_ks_dnum; r.perm = out->perm; }
        const int32_t *pi = probe_idx ? probe_idx + lo : nullptr;
        std::vector<int32_t> local;
        if (m->sliced && pi) {
            /* global methylation row -> row of this device's slice */
            const std::vector<int32_t> &u = m->rows_of[(size_t)i];
            local.resize((size_t)(hi - lo));
            for (int64_t k = 0; k < hi - lo; k++) {
                if (pi[k] < 0) { local[(size_t)k] = -1; continue; }
                auto it = std::lower_bound(u.begin(), u.end(), pi[k]);
                if (it == u.end() || *it != pi[k]) {
                    if (i == 0 && m->ev_expr) expr_ready.store(-1, std::memory_order_release); /* the other devices wait for this one's upload */
                    return fail(e, el, EPX_ESTATE, "pair %lld needs methylation row %d, which is not in this device's slice "
                                "(the index differs from the one passed to epx_multi_set_methylation_sliced)",
                                (long long)(lo + k), pi[k]);
                }
                local[(size_t)k] = (int32_t)(it - u.begin());
            }
            pi = local.data();
        }
        /* epx_analyze, with the expression matrix taken from device 0 by every other device: W uploads of the same 78 MB
         * from one host buffer made a call on 8 GPUs slower than on one */
        int rc;
        const bool trace = getenv("EPX_TRACE") != nullptr; /* host-side phase times of every device thread, to stderr */
        auto t_prev = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (!trace) return;
            const auto t = std::chrono::steady_clock::now();
            fprintf(stderr, "[epx_multi_analyze] device %d %-16s %8.3f ms\n", s->device, what,
                    std::chrono::duration<double, std::milli>(t - t_prev).count());
            t_prev = t;
        };
        if (i == 0 || !m->ev_expr) {
            rc = set_expression_begin(s, expr, n_genes, n_samples, 0, e, el);
            if (i == 0 && m->ev_expr) {
                if (rc == EPX_OK && cudaEventRecord(m->ev_expr, s->stream) != cudaSuccess) rc = fail(e, el, EPX_ECUDA, "event record failed");
                expr_ready.store(rc == EPX_OK ? 1 : -1, std::memory_order_release);
            }
        } else {
            int st;
            while ((st = expr_ready.load(std::memory_order_acquire)) == 0) std::this_thread::yield();
            if (st < 0) return fail(e, el, EPX_ESTATE, "the expression upload on the first device failed");
            if (cudaSetDevice(s->device) != cudaSuccess || cudaStreamWaitEvent(s->stream, m->ev_expr, 0) != cudaSuccess)
                return fail(e, el, EPX_ECUDA, "cannot wait for the first device's expression upload: %s", cudaGetErrorString(cudaGetLastError()));
            rc = set_expression_begin(s, m->sess[0]->expr.p, n_genes, n_samples, 1, e, el);
        }
        if (rc) return rc;
        lap("expression begin");
        rc = set_index_impl(s, rp.data(), pi, n_genes, meth_rows != nullptr, e, el);
        if (rc) { cudaStreamSynchronize(s->stream); return rc; }
        lap("index");
        rc = set_expression_finish(s, n_genes, n_samples, e, el);
        if (rc) return rc;
        lap("expression there");
        if (meth_rows) {
            const double *src = rows_readable_by_device(meth_rows, n_probes, n_samples, ld); /* as THIS device sees it */
            if (!src) return fail(e, el, EPX_EINVAL, "the methylation matrix is not page-locked for this device (use epx_host_alloc: portable)");
            rc = host_rows_fetch(s, src, ld, e, el);
            if (rc) return rc;
            lap("rows fetched");
        }
        rc = epx_session_run(s, opt, nullptr, e, el);
        if (rc) return rc;
        lap("run (enqueue)");
        rc = epx_session_fetch(s, &r, nullptr, e, el);
        lap("fetch");
        if (rc == EPX_OK && i == 0) { memcpy(out->counts, r.counts, sizeof r.counts); out->ref_gene = r.ref_gene; }
        return rc;
    }, err, errlen);
    if (rc_all == EPX_OK)
        for (int i = 0; i <= W; i++) m->pair_lo.push_back(row_ptr[cut[(size_t)i]]);
    return rc_all;
}

int epx_multi_analyze(epx_multi *m, const double *expr, int64_t n_genes, int64_t n_samples, const int64_t *row_ptr,
                      const int32_t *probe_idx, const epx_options *opt, epx_result *out, char *err, size_t errlen)
{
    return multi_analyze_impl(m, nullptr, 0, 0, expr, n_genes, n_samples, row_ptr, probe_idx, opt, out, err, errlen);
}

int epx_multi_analyze_host(epx_multi *m, const double *meth_rows, int64_t n_probes, int64_t ld, const double *expr,
                           int64_t n_genes, int64_t n_samples, const int64_t *row_ptr, const int32_t *probe_idx,
                           const epx_options *opt, epx_result *out, char *err, size_t errlen)
{
    if (!m || m->sess.empty() || !meth_rows || n_probes < 1 || ld < n_samples) return fail(err, errlen, EPX_EINVAL, "bad methylation arguments");
    int rc = check_samples(m->sess[0], n_samples, err, errlen);
    if (rc) return rc;
    EPX_CUDA(cudaSetDevice(m->sess[0]->device));
    if (getenv("EPX_NO_ZEROCOPY") || !rows_readable_by_device(meth_rows, n_probes, n_samples, ld)) {
        /* pageable memory: replicas of the whole matrix through the staged upload, then the resident-matrix call */
        rc = epx_multi_set_methylation_rows(m, meth_rows, n_probes, n_samples, ld, err, errlen);
        if (rc) return rc;
        return epx_multi_analyze(m, expr, n_genes, n_samples, row_ptr, probe_idx, opt, out, err, errlen);
    }
    m->sliced = false;      /* every device indexes its (partial) matrix by the caller's row numbers */
    m->rows_of.clear();
    return multi_analyze_impl(m, meth_rows, n_probes, ld, expr, n_genes, n_samples, row_ptr, probe_idx, opt, out, err, errlen);
}

int epx_multi_groups(epx_multi *m, const int64_t *group_ptr, const int32_t *group_pairs, int64_t n_groups,
                     epx_group_result *out, char *err, size_t errlen)
{
    if (!m || !group_ptr || !out || n_groups < 0) return fail(err, errlen, EPX_EINVAL, "bad group arguments");
    if (m->pair_lo.empty()) return fail(err, errlen, EPX_ESTATE, "epx_multi_groups before epx_multi_analyze");
    if (group_ptr[0] != 0) return fail(err, errlen, EPX_EINVAL, "bad group_ptr");
    const int W = (int)m->sess.size();
    /* a group lies inside one gene, hence on one device: the one that owns its first pair (empty groups: device 0) */
    std::vector<std::vector<int64_t>> mine((size_t)W);
    for (int64_t q = 0; q < n_groups; q++) {
        const int64_t b = group_ptr[q], e = group_ptr[q + 1];
        if (e < b) return fail(err, errlen, EPX_EINVAL, "group_ptr is not monotone at group %lld", (long long)q);
        int dev = 0;
        if (e > b) {
            if (!group_pairs) return fail(err, errlen, EPX_EINVAL, "group_pairs is NULL");
            const int64_t first = group_pairs[b];
            if (first < 0 || first >= m->pair_lo[(size_t)W])
                return fail(err, errlen, EPX_EINVAL, "group %lld names pair %lld outside [0,%lld)", (long long)q, (long long)first, (long long)m->pair_lo[(size_t)W]);
            dev = (int)(std::upper_bound(m->pair_lo.begin(), m->pair_lo.end(), first) - m->pair_lo.begin()) - 1;
            if (dev >= W) dev = W - 1;
            for (int64_t j = b; j < e; j++)
                if (group_pairs[j] < m->pair_lo[(size_t)dev] || group_pairs[j] >= m->pair_lo[(size_t)dev + 1])
                    return fail(err, errlen, EPX_EINVAL, "group %lld mixes genes of two devices", (long long)q);
        }
        mine[(size_t)dev].push_back(q);
    }
    return for_each_device(m, [&](int i, char *e, size_t el) -> int {
        const std::vector<int64_t> &qs = mine[(size_t)i];
        if (qs.empty()) return EPX_OK;
        const size_t Q = qs.size();
        const int64_t lo = m->pair_lo[(size_t)i];
        std::vector<int64_t> ptr(Q + 1, 0);
        std::vector<int32_t> pairs;
        for (size_t k = 0; k < Q; k++) {
            for (int64_t j = group_ptr[qs[k]]; j < group_ptr[qs[k] + 1]; j++) pairs.push_back((int32_t)(group_pairs[j] - lo));
            ptr[k + 1] = (int64_t)pairs.size();
        }
        std::vector<double> g(out->group ? Q * PAIR_COLS : 0), st(out->group_stat ? Q * 3 : 0);
        std::vector<uint16_t> na(out->group_na ? Q : 0);
        std::vector<int64_t> dn(out->group_ks_dnum ? Q * 3 : 0);
        epx_group_result r;
        r.group = out->group ? g.data() : nullptr; r.group_na = out->group_na ? na.data() : nullptr;
        r.group_stat = out->group_stat ? st.data() : nullptr; r.group_ks_dnum = out->group_ks_dnum ? dn.data() : nullptr;
        epx_session *s = m->sess[(size_t)i];
        int rc = epx_session_run_groups(s, ptr.data(), pairs.data(), (int64_t)Q, nullptr, e, el);
        if (rc == EPX_OK) rc = epx_session_fetch_groups(s, &r, nullptr, e, el);
        if (rc != EPX_OK) return rc;
        for (size_t k = 0; k < Q; k++) { /* rows of different devices are disjoint: no locking */
            const size_t q = (size_t)qs[k];
            if (out->group) memcpy(out->group + q * PAIR_COLS, g.data() + k * PAIR_COLS, PAIR_COLS * sizeof(double));
            if (out->group_na) out->group_na[q] = na[k];
            if (out->group_stat) memcpy(out->group_stat + q * 3, st.data() + k * 3, 3 * sizeof(double));
            if (out->group_ks_dnum) memcpy(out->group_ks_dnum + q * 3, dn.data() + k * 3, 3 * sizeof(int64_t));
        }
        return EPX_OK;
    }, err, errlen);
}

} /* extern "C" */
