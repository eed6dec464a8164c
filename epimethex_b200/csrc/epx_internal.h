/*
 * epx_internal.h — kernel launch interface between epx_api.cu (session, C ABI) and epx_kernels.cu.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/epx.h"

namespace epx {

/* work item of the pair kernel: a run of consecutive pairs of one gene */
struct Item {
    int32_t gene;
    int32_t begin;  /* first pair (index into probe_idx) */
    int32_t count;
    int32_t pad;
};

constexpr int PAIR_COLS = 11;
constexpr int GENE_COLS = 15;
constexpr int ITEM_PAIRS = 32;       /* pairs per work item (one deferred p-value per lane) */
constexpr int MAX_WARP_SAMPLES = 1024; /* warp-per-pair kernel: n <= 32 lanes * 32 elements */

/* device status word (d_status) */
enum { ST_OK = 0, ST_BREAKS_NOT_UNIQUE = 1 };

struct GeneParams {
    const double *expr;     /* [G][n] */
    int n, G, npad, m;
    uint16_t *perm;         /* [G][n] */
    uint8_t *lab8;          /* [G][ldl]  methylation-side label of each SAMPLE: rank/m, 3 = unlabelled */
    int64_t ldl;
    double *xc;             /* [G][ldx]  x - mean(x), original sample order */
    int64_t ldx;
    double *gene_aux;       /* [G][4]    {sxx = sum (x-mean)^2 (0 for a constant gene), mean, sum (x-mean), 1/sqrt(sxx)} */
    int32_t *nuniq;         /* [G] */
    unsigned one;           /* = 1, opaque to the compiler (see ce_fma) */
};

struct GeneTestParams {
    const double *expr;
    const uint16_t *perm;
    int n, G;
    const int32_t *nuniq;
    int32_t *counts;        /* [3] nUP, nM, nD */
    int32_t *ref_gene;
    int32_t *status;
    int test_t, data_linear, fc_from_means;
    double *gene;           /* [G][15] */
    uint16_t *gene_na;
    int32_t *gene_dnum;     /* [G][3] */
};

struct RankParams {
    const double *meth;
    int64_t ld;
    const int32_t *uniq_probe; /* [U] */
    int64_t U;
    int n, npad;
    uint16_t *ord;          /* [U][ldo]: sample index of the k-th smallest value | 0x8000 when equal to the next */
    int64_t ldo;
    unsigned one;           /* = 1, opaque to the compiler (see ce_fma) */
    int32_t *queue;         /* k_probe_rank_w: {next ticket, warps done}, both 0 between launches */
    int32_t *gene_queue;    /* the same for its expression-row instantiation (the two launches run side by side) */
    /* k_probe_rank_w<EPS, true> (launch_gene_rows) orders the expression rows with the same loop and its last warp
     * picks the bin.var reference gene: the gene-level work of R/EpimethexAnalysis.R:117-139 */
    GeneParams genes;
    const int32_t *nuniq_in; /* = genes.nuniq */
    int32_t *counts, *ref_gene, *status;
    unsigned long long *best;   /* max over genes of (distinct values << 32 | ~gene); 0 between launches */
    int32_t *full_rows;         /* rows whose bucket placement failed its check and went through the full network (may be NULL) */
};

struct PairParams {
    const double *meth;
    int64_t ld;
    const uint16_t *ord;
    int64_t ldo;
    const int32_t *probe_idx;
    const int32_t *pair_slot;
    const double *xc;
    int64_t ldx;
    const uint8_t *lab8;
    int64_t ldl;
    const uint16_t *perm;
    const double *gene_aux;
    const Item *items;
    int n_items;
    int32_t *queue;           /* k_pair_stats2: {next ticket, warps done}, both 0 between launches */
    const int32_t *pair_gene; /* [Np] gene of each pair (CTA-per-pair kernel) */
    int pair_begin, n_pairs;  /* CTA-per-pair kernel: pairs [pair_begin, n_pairs) */
    int n, m;
    int test_t;
    int exact_ok;           /* m*m < 10000: exact KS p-value unless the two groups hold ties */
    unsigned warp_smem;     /* k_pair_stats2: bytes of shared memory per warp */
    const double *pt_tab;   /* [pt_n + 1] epx_pearson_h on the grid r = i / pt_n (see epx_math.h) */
    int pt_n;
    /* t mode: Welch p-values by the (df, r) table of epx_welch_p_tab (epx_math.h); NULL = continued fraction (m < 30) */
    const double *wt_tab;
    int wt_N, wt_ld, wt_K;
    double wt_df0, wt_inv_ddf;
    const double *lut_exact; /* [m+1] p-value for max|cA-cB| = d */
    const double *lut_asym;  /* [m+1] */
    double *pair;           /* [Np][11] */
    uint16_t *pair_na;
    double *pair_stat;      /* [Np][3] */
    int32_t *pair_dnum;     /* [Np][3] */
};

/* ---- pooled probe groups (epx_group.cuh) ---- */
struct GroupDesc {
    int32_t gene;
    int32_t k;        /* probes (all with data) */
    int32_t first;    /* offset of the group's pair list in group_pairs */
    int32_t nblocks;  /* sorted runs written by k_group_blocksort */
    int64_t elem;     /* offset of the group's k*n stacked values in the work buffers */
};
struct BlockDesc { int32_t group, r0, runs, pad; };   /* probes [r0, r0 + runs) of a group */
struct TileDesc { int32_t group, tile; };             /* MERGE_TILE outputs of one group */

struct GroupParams {
    const double *meth;
    int64_t ld;
    const int32_t *probe_idx;
    const int32_t *group_pairs;
    const int32_t *pair_slot;   /* order row of each pair's probe (k_probe_rank) */
    const uint16_t *ord;
    int64_t ldo;
    const GroupDesc *groups;
    const BlockDesc *blocks;
    const TileDesc *tiles;
    const int32_t *multi;       /* groups handled by k_group_moments / k_group_walk: more than one block, or empty */
    int n_multi;
    int n_groups, n_blocks;
    int n, m, cap, test_t;
    const uint16_t *perm;
    const uint8_t *lab8;
    int64_t ldl;
    const double *xc;
    int64_t ldx;
    const double *gene_aux;
    double *val[2];         /* ping-pong: stacked values of every group, ascending inside sorted runs */
    uint8_t *lab[2];        /* their labels */
    int32_t *guard;         /* [n_groups][3] calculateTtest guard: 1 differs, 0 all equal, -1 NA */
    double *grp;            /* [n_groups][11] */
    uint16_t *grp_na;
    double *grp_stat;       /* [n_groups][3] */
    int64_t *grp_dnum;      /* [n_groups][3] */
    unsigned one;
};

cudaError_t launch_group_moments(const GroupParams &p, cudaStream_t st);
cudaError_t launch_group_blocksort(const GroupParams &p, cudaStream_t st);
cudaError_t launch_group_merge(const GroupParams &p, int pass, int64_t run_len, int n_tiles, cudaStream_t st);
cudaError_t launch_group_walk(const GroupParams &p, cudaStream_t st);
int group_block_cap(int n);      /* entries of one k_group_blocksort block */

cudaError_t launch_gene_sort(const GeneParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_pick_reference(const GeneTestParams &p, cudaStream_t st);
cudaError_t launch_gene_tests(const GeneTestParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_probe_rank(const RankParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_gene_rows(const RankParams &p, int sm_count, cudaStream_t st);
cudaError_t launch_pair_stats(const PairParams &p, int sm_count, cudaStream_t st);
/* sample counts whose expression rows go through launch_gene_rows (else launch_gene_sort + launch_pick_reference) */
bool probe_rank_takes_genes(int n);

/* dst[p*ld + s0 + j] = src[j*P + p], j < ns : sample-major chunk -> probe-major rows */
/* every upload kernel records in *first_bad (atomicMin) the smallest row * n + sample whose value is NaN or infinite */
cudaError_t launch_transpose_cols(const double *src, double *dst, int64_t P, int ns, int64_t s0, int64_t ld,
                                  int64_t n_all, unsigned long long *first_bad, cudaStream_t st);
/* dst[r*ld + j] = src[r*n + j]: re-space contiguous rows to the padded leading dimension (row0 = first row's number) */
cudaError_t launch_respace_rows(const double *src, double *dst, int64_t rows, int64_t n, int64_t ld, int64_t row0,
                                unsigned long long *first_bad, cudaStream_t st);
/* dst[p*ld + j] = src[p*lds + j] for the rows p = uniq[0..U) only; src is HOST memory the device can read in place
 * (page-locked), fetched over PCIe by the kernel itself (epx_analyze_host) */
cudaError_t launch_gather_host_rows(const double *src, int64_t lds, const int32_t *uniq, int64_t U, int64_t n, double *dst,
                                    int64_t ld, unsigned long long *first_bad, cudaStream_t st);
/* finite test alone, for matrices that need no re-spacing */
cudaError_t launch_check_finite(const double *a, int64_t rows, int64_t n, int64_t ld, unsigned long long *first_bad,
                                cudaStream_t st);
/* one-time kernel attribute setup (dynamic shared memory opt-in) */
cudaError_t kernels_init();

} // namespace epx
