// ft8_b200.cu -- handle, table upload and the C ABI (include/ft8_b200.h) of the B200-native FT8 receive path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see __graft_entry__.build).
// One handle = one device + one stream + constant tables + scratch sized from ft8_cfg.max_cycles.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ft8_b200.h"
#include "ft8_tables.h"
#include "fft.cuh"
#include "codec.cuh"
#include "ldpc.cuh"
#include "osd.cuh"
#include "spectrogram.cuh"
#include "sync.cuh"
#include "fine.cuh"
#include "fine_tc.cuh"
#include "passes.cuh"
#include "synth.cuh"

using namespace ft8;

static thread_local std::string g_create_error;

struct ft8_handle {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // host->device audio copies, overlapped with S1/S2/F1 of earlier chunks
    cudaEvent_t chunk_ev[16] = {};
    ft8_cfg cfg{};
    std::string err;
    // constant-like tables in global memory
    float* d_hann = nullptr;
    float2 *d_W1920 = nullptr, *d_W3840 = nullptr, *d_W3200 = nullptr, *d_W375 = nullptr, *d_W256 = nullptr,
           *d_W96000 = nullptr, *d_W192000 = nullptr, *d_W32 = nullptr;
    float2 *d_TS = nullptr, *d_TF = nullptr, *d_TC = nullptr, *d_T256 = nullptr, *d_W96000T = nullptr;   // per-pass twiddle tables [k-1][p]
    float* d_pulse = nullptr;        // GFSK pulse for the synthetic generator
    // tensor-core frequency scan (fine_tc.cuh): B operand (E | M, tf32 hi/lo), half-sample phase table, per-item intermediates
    float* d_bmat = nullptr; float2* d_w6400 = nullptr;
    float2* d_zwin = nullptr; TScanOut* d_tso = nullptr; int32_t* d_ff = nullptr; size_t fine_tmp_items = 0;
    int32_t* d_amb = nullptr;         // [fine_tmp_items + 1]: near-tie candidates of the frequency scan ([0] = count, list from [1])
    // batch scratch
    size_t cap_cycles = 0, cap_slots = 0;
    void* d_audio = nullptr; size_t audio_bytes = 0;
    // prefetch slot (ft8_prefetch_audio): the NEXT batch's audio is copied here while the current batch is decoded
    void* d_audio_pf = nullptr; size_t audio_pf_bytes = 0;
    const void* pf_host = nullptr; int pf_B = 0, pf_dtype = -1;
    cudaEvent_t pf_ev = nullptr;
    float* d_grid = nullptr;
    // live mode (ft8_decode_cycles_live): the reference's two-cycle waterfall ring per stream + the previous cycle's last window
    float* d_ring = nullptr; void* d_tail = nullptr; int tail_dtype = -1; bool tail_valid = false;
    float2* d_Y = nullptr; size_t y_cycles = 0;
    float2* d_spec = nullptr;
    float* d_best_score = nullptr; int16_t* d_best_h0 = nullptr;
    int16_t *d_f0 = nullptr, *d_h0 = nullptr; float* d_score = nullptr; int32_t* d_ncand = nullptr;
    int32_t* d_cycle_of = nullptr;
    uint8_t* d_status = nullptr; float* d_llr_grid = nullptr; float* d_grid_sd = nullptr; int8_t* d_grid_snr = nullptr;
    float* d_llr_fine = nullptr; FineOut* d_fine = nullptr; float* d_saved = nullptr; uint8_t* d_saved_n = nullptr;
    uint8_t* d_saved_ap = nullptr; uint32_t* d_bits = nullptr; uint8_t *d_ripass = nullptr, *d_rap = nullptr, *d_rmethod = nullptr;
    uint16_t* d_rnits = nullptr; int32_t* d_osd_found = nullptr; uint32_t* d_osd_bits = nullptr;
    int32_t *d_list_fine = nullptr, *d_list_osd = nullptr, *d_counts = nullptr;   // counts: [0] fine, [1] osd, [2] records
    DevStats* d_stats = nullptr;
    ft8_record* d_rec = nullptr;
    int32_t *d_rec_n = nullptr, *d_rec_base = nullptr;   // per-cycle record counts / offsets
    int32_t* h_counts = nullptr;      // pinned
    DevStats* h_stats = nullptr;      // pinned
    // generic arena for the stand-alone stage ops
    void* arena = nullptr; size_t arena_bytes = 0;
    cudaEvent_t ev[14] = {};          // ev[0..8]: stage boundaries of ft8_decode_cycles; ev[10], ev[11]: stand-alone ops;
                                      // ev[12], ev[13]: between the three kernels of the fine stage
    float last_ms[12] = {};           // see ft8_last_kernel_ms
    ft8_stats stats{};
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return FT8_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

static int fail(ft8_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

// ------------------------------------------------------------------------------------------ tables
static std::vector<float2> twiddles(int n, int count) {
    std::vector<float2> w(count);
    for (int j = 0; j < count; ++j) {
        const double a = -2.0 * M_PI * (double)j / (double)n;
        w[j] = make_float2((float)cos(a), (float)sin(a));
    }
    return w;
}

static float2 twiddle1(int n, long long j) {
    const double a = -2.0 * M_PI * (double)(j % n) / (double)n;
    return make_float2((float)cos(a), (float)sin(a));
}

// Per-pass twiddle table of Stockham pass (R, S) of a length-n transform, laid out [k-1][p] (fft.cuh, Pass::compute_store
// with TT = true): entry = w_n^(p k S), the same value the plain table holds at index p*k*S.
static void append_pass_table(std::vector<float2>& out, int n, int R, int S) {
    const int M = n / (R * S);
    for (int k = 1; k < R; ++k)
        for (int p = 0; p < M; ++p) out.push_back(twiddle1(n, (long long)p * k * S));
}

template <typename T> static cudaError_t upload(T** dst, const std::vector<T>& v) {
    cudaError_t e = cudaMalloc((void**)dst, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static uint32_t crc14_bits(const uint8_t* msg77) {      // decoders.py:123-129
    uint32_t r = 0;
    for (int i = 0; i < 96; ++i) {
        const uint32_t bit = (i < 77) ? msg77[i] : 0u;
        const uint32_t top = (r >> 13) & 1u;
        r = ((r << 1) & 0x3FFFu) | bit;
        if (top) r ^= 0x2757u;
    }
    return r;
}

static cudaError_t upload_constant_tables() {
    // CRC syndrome contributions
    CodecTables ct;
    memset(&ct, 0, sizeof(ct));
    for (int j = 0; j < 91; ++j) {
        if (j < 77) {
            uint8_t m[77] = {0};
            m[j] = 1;
            ct.crc_syn[j] = (uint16_t)crc14_bits(m);
        } else {
            ct.crc_syn[j] = (uint16_t)(1u << (90 - j));
        }
    }
    memcpy(ct.prefix2, FT8_PREFIX2_BITS, sizeof(ct.prefix2));
    cudaError_t e = cudaMemcpyToSymbol(c_codec, &ct, sizeof(ct));
    if (e != cudaSuccess) return e;
    // LDPC graph; variable -> edge slots in np.add.at order (degree-6 group flattened first, then degree-7)
    LdpcTables lt;
    memset(&lt, 0, sizeof(lt));
    int cnt[174] = {0};
    for (int c = 0; c < 83; ++c)
        for (int k = 0; k < 7; ++k) {
            const uint8_t v = FT8_CHECK_VARS[c][k];
            lt.chk_var[c * 7 + k] = v;
            if (v != 255) lt.var_edge[3 * v + cnt[v]++] = (uint16_t)(c * 7 + k);
        }
    e = cudaMemcpyToSymbol(c_ldpc, &lt, sizeof(lt));
    if (e != cudaSuccess) return e;
    // OSD columns of G0 = [I | A^T]
    OsdTables ot;
    memset(&ot, 0, sizeof(ot));
    for (int c = 0; c < 91; ++c) ot.col[c][c >> 5] = 1u << (c & 31);
    for (int i = 0; i < 83; ++i)
        for (int w = 0; w < 3; ++w) ot.col[91 + i][w] = FT8_GEN_MASK[i][w];
    e = cudaMemcpyToSymbol(c_osd, &ot, sizeof(ot));
    if (e != cudaSuccess) return e;
    // AP patterns (receiver.py:21-27)
    ApTables ap;
    memset(&ap, 0, sizeof(ap));
    const char* pat[5] = {"", "00000000000000000000000000100", "0111111001110101001", "0111111010010100001", "0111111010010010001"};
    const int first[5] = {0, 0, 58, 58, 58};
    for (int a = 0; a < 5; ++a) {
        ap.first[a] = (int8_t)first[a];
        ap.len[a] = (int8_t)strlen(pat[a]);
        for (int i = 0; i < ap.len[a]; ++i) if (pat[a][i] == '1') ap.mask[a] |= 1u << i;
    }
    e = cudaMemcpyToSymbol(c_ap, &ap, sizeof(ap));
    if (e != cudaSuccess) return e;
    // fine-stage tables
    FineTables ft;
    for (int i = 0; i < 100; ++i) {
        const double x = -M_PI + M_PI * (double)i / 99.0;       // np.linspace(-pi, 0, 100); cos is even so the
        ft.taper[i] = (float)(0.5 * (1.0 + cos(x)));             // (pi, 0) variant of receiver.py:183 is identical
    }
    for (int m = 0; m < 32; ++m) {
        const double a = -2.0 * M_PI * m / 32.0;
        ft.w32[m] = make_float2((float)cos(a), (float)sin(a));
    }
    return cudaMemcpyToSymbol(c_fine, &ft, sizeof(ft));
}

// B operand of k_fscan_mma, laid out exactly as one shared-memory stage per K-chunk: [chunk][split hi/lo][k-chunk 0/1][n][4]
// (see fine_tc.cuh for the formulas; tp is the reference's taper incl. the inverted upper ramp, receiver.py:182-183)
static std::vector<float> fscan_bmat() {
    auto taper = [](int i) { return 0.5 * (1.0 + cos(-M_PI + M_PI * (double)i / 99.0)); };
    auto tp = [&](int k) -> double {
        if (k < -150 || k >= 850) return 0.0;
        if (k < -50) return taper(k + 150);
        if (k < 750) return 1.0;
        return taper(k - 750);
    };
    auto split_hi = [](float v) { uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    std::vector<float> out((size_t)FS_BMAT_FLOATS, 0.0f);
    for (int n = 0; n < FS_N; ++n) {
        const int f = n >> 3, t = n & 7;
        const int d = -32 + 8 * (f < 4 ? f : f + 1);
        for (int k = 0; k < FS_K; ++k) {
            double v;
            if (k < FS_K1) {
                const int r = k >> 1;
                const double ang = -2.0 * M_PI * (double)(100 * t + d) * ((double)r - 15.5) / 3200.0;
                v = (k & 1) ? sin(ang) : cos(ang);
            } else {
                const int u = fs_u_of(k - FS_K1);
                const int x = u - d - 100 * t;
                const double dir = x == 0 ? 32.0 : sin(M_PI * x / 100.0) / sin(M_PI * x / 3200.0);
                v = (tp(u - d) - tp(u)) * dir;
            }
            const float vf = (float)v, hi = split_hi(vf), lo = split_hi((float)(v - (double)hi));
            const int chunk = k / FS_KC, kc = (k % FS_KC) / 4, e = k % 4;
            const size_t base = (size_t)chunk * (FS_B_STAGE_BYTES / 4);
            out[base + 0 * (FS_B_SPLIT_BYTES / 4) + kc * (FS_N * 4) + n * 4 + e] = hi;
            out[base + 1 * (FS_B_SPLIT_BYTES / 4) + kc * (FS_N * 4) + n * 4 + e] = lo;
        }
    }
    return out;
}

template <typename T> static cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

static int ensure_arena(ft8_handle* h, size_t bytes) {
    if (bytes <= h->arena_bytes) return FT8_OK;
    if (h->arena) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->arena)); h->arena = nullptr; h->arena_bytes = 0; }
    CK(cudaMalloc(&h->arena, bytes));
    h->arena_bytes = bytes;
    return FT8_OK;
}

static int ensure_audio(ft8_handle* h, size_t bytes) {
    if (bytes <= h->audio_bytes) return FT8_OK;
    if (h->d_audio) { CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->d_audio)); h->d_audio = nullptr; h->audio_bytes = 0; }
    CK(cudaMalloc(&h->d_audio, bytes));
    h->audio_bytes = bytes;
    return FT8_OK;
}

extern "C" void ft8_default_cfg(ft8_cfg* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->max_cycles = 1;
    cfg->max_cands = 200;
    cfg->sync_score_min = 85.0f;
    cfg->llr_sd_min = 5.0f;
    cfg->osd_singleflips = 30;
    cfg->osd_doubleflips = 2;
    cfg->max_codewords = 1 << 16;
}

extern "C" int ft8_create(int device, const ft8_cfg* cfg_in, ft8_handle** out) {
    if (!out) return fail(nullptr, FT8_E_BADARG, "ft8_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, FT8_E_NODEVICE, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, FT8_E_BADARG, "ft8_create: bad device index");
    ft8_cfg cfg;
    ft8_default_cfg(&cfg);
    if (cfg_in) cfg = *cfg_in;
    if (cfg.max_cycles <= 0) cfg.max_cycles = 1;
    if (cfg.max_cands <= 0) cfg.max_cands = 200;
    if (cfg.max_cands > N_F0) cfg.max_cands = N_F0;
    if (cfg.osd_singleflips < 0 || cfg.osd_singleflips > OSD_MAX_FLIPS || cfg.osd_doubleflips < 0)
        return fail(nullptr, FT8_E_BADARG, "ft8_create: osd flips out of range");
    if (cfg.max_codewords <= 0) cfg.max_codewords = 1 << 16;
    if (cfg.search_f0_lo == 0 && cfg.search_f0_hi == 0) { cfg.search_f0_lo = F0_LO; cfg.search_f0_hi = F0_LO + N_F0; }
    if (cfg.search_h0_lo == 0 && cfg.search_h0_hi == 0) { cfg.search_h0_lo = H0_LO; cfg.search_h0_hi = H0_LO + N_H0; }
    if (cfg.search_f0_lo < F0_LO || cfg.search_f0_hi > F0_LO + N_F0 || cfg.search_f0_lo >= cfg.search_f0_hi ||
        cfg.search_h0_lo < H0_LO || cfg.search_h0_hi > H0_LO + N_H0 || cfg.search_h0_lo >= cfg.search_h0_hi)
        return fail(nullptr, FT8_E_BADARG, "ft8_create: search range outside what the kernels are built for (f0 in [32, 960), h0 in [-37, 87))");
    if (cfg.fine_mode != 0 && cfg.fine_mode != 1) return fail(nullptr, FT8_E_BADARG, "ft8_create: fine_mode must be 0 (tensor-core scan) or 1 (FFT scan)");
    ft8_handle* h = new ft8_handle();
    h->device = device;
    h->cfg = cfg;
#define CKC(call)                                                                              \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e__);              \
            ft8_destroy(h);                                                                    \
            return FT8_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    h->n_sm = prop.multiProcessorCount;
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : h->ev) CKC(cudaEventCreate(&ev));
    for (auto& ev : h->chunk_ev) CKC(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->pf_ev, cudaEventDisableTiming));
    CKC(upload_constant_tables());
    {
        std::vector<float> w(NFFT_S);
        for (int i = 0; i < NFFT_S; ++i) {          // np.hanning: 0.5 + 0.5*cos(pi*n/(M-1)), n = 1-M, 3-M, ...
            const double n = (double)(1 - NFFT_S + 2 * i);
            w[i] = (float)(0.5 + 0.5 * cos(M_PI * n / (double)(NFFT_S - 1)));
        }
        CKC(upload(&h->d_hann, w));
        CKC(upload(&h->d_W1920, twiddles(1920, 1920)));
        CKC(upload(&h->d_W3840, twiddles(3840, 1921)));
        CKC(upload(&h->d_W3200, twiddles(3200, 3200)));
        CKC(upload(&h->d_W375, twiddles(375, 375)));
        CKC(upload(&h->d_W256, twiddles(256, 256)));
        CKC(upload(&h->d_W192000, twiddles(192000, 96001)));
        CKC(upload(&h->d_W32, twiddles(32, 32)));
        {   // per-pass tables (coalesced twiddle loads), one buffer per kernel
            std::vector<float2> ts, tf, tc, t256, w96t;
            append_pass_table(ts, 1920, 15, 1); append_pass_table(ts, 1920, 8, 15);                              // SP_T8_OFF
            append_pass_table(tf, 3200, 5, 1); append_pass_table(tf, 3200, 5, 5); append_pass_table(tf, 3200, 8, 25);   // FINE_T*_OFF
            append_pass_table(tc, 375, 3, 1); append_pass_table(tc, 375, 5, 3); append_pass_table(tc, 375, 5, 15);
            append_pass_table(t256, 256, 16, 1);
            w96t.reserve((size_t)CS_N1 * CS_N2);
            for (int k1 = 0; k1 < CS_N1; ++k1)
                for (int n2 = 0; n2 < CS_N2; ++n2) w96t.push_back(twiddle1(96000, (long long)n2 * k1));
            if ((int)ts.size() != SP_T8_OFF + 7 * 16 || (int)tf.size() != FINE_TF_LEN || (int)tc.size() != 370) { ft8_destroy(h); return fail(nullptr, FT8_E_CUDA, "twiddle table layout"); }
            CKC(upload(&h->d_TS, ts)); CKC(upload(&h->d_TF, tf)); CKC(upload(&h->d_TC, tc)); CKC(upload(&h->d_T256, t256));
            CKC(upload(&h->d_W96000T, w96t));
        }
        CKC(upload(&h->d_pulse, synth_pulse_table()));
        CKC(upload(&h->d_bmat, fscan_bmat()));
        {
            std::vector<float2> w(FS_W_LEN);
            for (int g = 0; g < FS_W_LEN; ++g) { const double a = 2.0 * M_PI * (double)g / (double)FS_W_LEN; w[g] = make_float2((float)cos(a), (float)sin(a)); }
            CKC(upload(&h->d_w6400, w));
        }
    }
    const size_t B = (size_t)cfg.max_cycles, K = (size_t)cfg.max_cands, N = B * K;
    h->cap_cycles = B;
    h->cap_slots = N;
    CKC(dmalloc(&h->d_grid, B * GRID_ROWS * GRID_COLS));
    // four-step scratch Y: 768 KB per cycle, written by k_cs_cols and read back by k_cs_rows, in chunks of up to 1024 cycles.
    // Smaller, L2-resident chunks were measured and are SLOWER: 3.18 ms per 4096 cycles at 1024, 3.50 at 256,
    // 3.88 at 128, 4.20 at 64, 5.00 at 32 (profiles/r02_experiments.md) -- both kernels sit on the L1/shared data pipe, not on
    // HBM, so the 3x DRAM traffic costs nothing while small grids lose to tail effects.  FT8_Y_CYCLES overrides for experiments.
    size_t ych = 1024;
    if (const char* e = getenv("FT8_Y_CYCLES")) ych = (size_t)std::max(1, atoi(e));
    h->y_cycles = std::min<size_t>(B, ych);
    CKC(dmalloc(&h->d_Y, h->y_cycles * CS_N));
    CKC(dmalloc(&h->d_spec, B * FINE_SPEC_STRIDE));
    CKC(dmalloc(&h->d_best_score, B * N_F0 * SY_HS));
    CKC(dmalloc(&h->d_best_h0, B * N_F0 * SY_HS));
    CKC(dmalloc(&h->d_f0, N)); CKC(dmalloc(&h->d_h0, N)); CKC(dmalloc(&h->d_score, N)); CKC(dmalloc(&h->d_ncand, B));
    CKC(dmalloc(&h->d_cycle_of, N));
    CKC(dmalloc(&h->d_status, N)); CKC(dmalloc(&h->d_llr_grid, N * 174)); CKC(dmalloc(&h->d_grid_sd, N)); CKC(dmalloc(&h->d_grid_snr, N));
    CKC(dmalloc(&h->d_llr_fine, N * 174)); CKC(dmalloc(&h->d_fine, N)); CKC(dmalloc(&h->d_saved, N * 5 * 174));
    CKC(dmalloc(&h->d_saved_n, N)); CKC(dmalloc(&h->d_saved_ap, N * 5)); CKC(dmalloc(&h->d_bits, N * 3));
    CKC(dmalloc(&h->d_ripass, N)); CKC(dmalloc(&h->d_rap, N)); CKC(dmalloc(&h->d_rmethod, N)); CKC(dmalloc(&h->d_rnits, N));
    CKC(dmalloc(&h->d_osd_found, N * 10)); CKC(dmalloc(&h->d_osd_bits, N * 30));
    CKC(dmalloc(&h->d_list_fine, N)); CKC(dmalloc(&h->d_list_osd, N)); CKC(dmalloc(&h->d_counts, 8));
    CKC(dmalloc(&h->d_stats, 1)); CKC(dmalloc(&h->d_rec, N)); CKC(dmalloc(&h->d_rec_n, B)); CKC(dmalloc(&h->d_rec_base, B));
    CKC(cudaMallocHost((void**)&h->h_counts, 8 * sizeof(int32_t)));
    CKC(cudaMallocHost((void**)&h->h_stats, sizeof(DevStats)));
    {
        std::vector<int32_t> co(N);
        for (size_t i = 0; i < N; ++i) co[i] = (int32_t)(i / K);
        CKC(cudaMemcpy(h->d_cycle_of, co.data(), N * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    CKC(cudaFuncSetAttribute(k_fine, cudaFuncAttributeMaxDynamicSharedMemorySize, FINE_SMEM_BYTES));
    CKC(cudaFuncSetAttribute(k_fine_tscan<FT_NT, FT_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
    CKC(cudaFuncSetAttribute(k_fine_final<FT_NT, FF_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_BYTES));
    CKC(cudaFuncSetAttribute(k_fscan_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM_BYTES));
    if (cfg.fine_mode == 0) {
        h->fine_tmp_items = N;
        CKC(dmalloc(&h->d_zwin, N * FS_WIN)); CKC(dmalloc(&h->d_tso, N)); CKC(dmalloc(&h->d_ff, N)); CKC(dmalloc(&h->d_amb, N + 1));
    }
    {
        const int sp_smem = SP_ROWS * SP_BUFS * SP_BUF_LEN * (int)sizeof(float2);
        CKC(cudaFuncSetAttribute(k_spectrogram<int16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp_smem));
        CKC(cudaFuncSetAttribute(k_spectrogram<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp_smem));
        CKC(cudaFuncSetAttribute(k_spectrogram<int16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp_smem));
        CKC(cudaFuncSetAttribute(k_spectrogram<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sp_smem));
    }
    CKC(cudaFuncSetAttribute(k_sync_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM_BYTES));
    CKC(cudaDeviceSynchronize());
#undef CKC
    *out = h;
    return FT8_OK;
}

extern "C" void ft8_destroy(ft8_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* ptrs[] = {h->d_TS, h->d_TF, h->d_TC, h->d_T256, h->d_W96000T, h->d_hann, h->d_W1920, h->d_W3840, h->d_W3200, h->d_W375, h->d_W256, h->d_W96000, h->d_W192000, h->d_W32,
                    h->d_pulse, h->d_ring, h->d_tail, h->d_bmat, h->d_w6400, h->d_zwin, h->d_tso, h->d_ff, h->d_amb, h->d_audio, h->d_grid, h->d_Y, h->d_spec, h->d_best_score, h->d_best_h0, h->d_f0, h->d_h0, h->d_score,
                    h->d_ncand, h->d_cycle_of, h->d_status, h->d_llr_grid, h->d_grid_sd, h->d_grid_snr, h->d_llr_fine, h->d_fine,
                    h->d_saved, h->d_saved_n, h->d_saved_ap, h->d_bits, h->d_ripass, h->d_rap, h->d_rmethod, h->d_rnits,
                    h->d_osd_found, h->d_osd_bits, h->d_list_fine, h->d_list_osd, h->d_counts, h->d_stats, h->d_rec, h->d_rec_n, h->d_rec_base, h->arena};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->chunk_ev) if (ev) cudaEventDestroy(ev);
    if (h->pf_ev) cudaEventDestroy(h->pf_ev);
    if (h->d_audio_pf) cudaFree(h->d_audio_pf);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char* ft8_last_error(ft8_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" void* ft8_stream(ft8_handle* h) { return h ? (void*)h->stream : nullptr; }
extern "C" int ft8_host_alloc(size_t bytes, int flags, void** out) {
    if (!out || bytes == 0) return FT8_E_BADARG;
    *out = nullptr;
    unsigned f = cudaHostAllocPortable;
    if (flags & FT8_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    if (cudaHostAlloc(out, bytes, f) != cudaSuccess) { cudaGetLastError(); *out = nullptr; return FT8_E_CUDA; }
    return FT8_OK;
}
extern "C" int ft8_host_free(void* p) {
    if (!p) return FT8_OK;
    if (cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return FT8_E_CUDA; }
    return FT8_OK;
}
extern "C" int ft8_synchronize(ft8_handle* h) {
    if (!h) return FT8_E_BADARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}
extern "C" int ft8_get_stats(ft8_handle* h, ft8_stats* out) {
    if (!h || !out) return FT8_E_BADARG;
    *out = h->stats;
    return FT8_OK;
}
extern "C" int ft8_last_kernel_ms(ft8_handle* h, int which, float* ms) {
    if (!h || !ms || which < 0 || which > 11) return FT8_E_BADARG;
    *ms = h->last_ms[which];
    return FT8_OK;
}

// ------------------------------------------------------------------------------------------ launch helpers
static CandState cand_state(ft8_handle* h) {
    CandState cs;
    cs.K = h->cfg.max_cands; cs.n_cand = h->d_ncand; cs.f0 = h->d_f0; cs.h0 = h->d_h0; cs.score = h->d_score;
    cs.status = h->d_status; cs.llr_grid = h->d_llr_grid; cs.grid_sd = h->d_grid_sd; cs.grid_snr = h->d_grid_snr;
    cs.llr_fine = h->d_llr_fine; cs.fine = h->d_fine; cs.saved_llr = h->d_saved; cs.saved_n = h->d_saved_n;
    cs.saved_ap = h->d_saved_ap; cs.bits91 = h->d_bits; cs.r_ipass = h->d_ripass; cs.r_ap = h->d_rap;
    cs.r_method = h->d_rmethod; cs.r_nits = h->d_rnits; cs.osd_found = h->d_osd_found; cs.osd_bits = h->d_osd_bits;
    return cs;
}

static int launch_spectrogram(ft8_handle* h, const void* d_audio, int dtype, int B, float* d_grid, int row_lo = 1,
                              int row_hi = 375, int out_rows = GRID_ROWS, int out_row0 = 0, int fill_row0 = 1,
                              const void* d_prev_tail = nullptr, int out_wrap = 0) {
    const bool ring = d_prev_tail != nullptr || out_wrap != 0;
    dim3 grid((row_hi - row_lo + SP_ROWS) / SP_ROWS, B);
    const int smem = SP_ROWS * SP_BUFS * SP_BUF_LEN * (int)sizeof(float2);
#define SP_LAUNCH(T, RING)                                                                                                          \
    k_spectrogram<T, RING><<<grid, SP_ROWS * SP_NT, smem, h->stream>>>((const T*)d_audio, d_grid, h->d_hann, h->d_TS, h->d_W3840, row_lo, \
                                                                       row_hi, out_rows, out_row0, fill_row0, (const T*)d_prev_tail, out_wrap)
    if (dtype == FT8_AUDIO_I16) { if (ring) SP_LAUNCH(int16_t, true); else SP_LAUNCH(int16_t, false); }
    else { if (ring) SP_LAUNCH(float, true); else SP_LAUNCH(float, false); }
#undef SP_LAUNCH
    CK(cudaGetLastError());
    return FT8_OK;
}

// d_grid points at the first of the B cycles; b0 = index of that cycle in the handle's per-cycle arrays
static int launch_sync(ft8_handle* h, const float* d_grid, int grid_rows, int B, int odd_even, int b0 = 0) {
    const int cycle_h0 = odd_even ? 375 : 0;
    const size_t K = h->cfg.max_cands;
    k_sync_scores<<<dim3(SY_HS * (N_F0 / SY_TF), B), SY_NT, SY_SMEM_BYTES, h->stream>>>(
        d_grid, grid_rows, cycle_h0, h->d_best_score + (size_t)b0 * N_F0 * SY_HS, h->d_best_h0 + (size_t)b0 * N_F0 * SY_HS,
        h->cfg.search_h0_lo, h->cfg.search_h0_hi);
    CK(cudaGetLastError());
    k_topk<<<B, 960, 0, h->stream>>>(h->d_best_score + (size_t)b0 * N_F0 * SY_HS, h->d_best_h0 + (size_t)b0 * N_F0 * SY_HS, h->cfg.sync_score_min,
                                     h->cfg.max_cands, h->d_f0 + b0 * K, h->d_h0 + b0 * K, h->d_score + b0 * K, h->d_ncand + b0,
                                     h->cfg.search_f0_lo, h->cfg.search_f0_hi);
    CK(cudaGetLastError());
    return FT8_OK;
}

// cycle spectra for B cycles into spec[B][stride], processed in chunks bounded by the Y scratch
static int launch_cycle_spectrum(ft8_handle* h, const void* d_audio, int dtype, int B, float2* d_spec, int stride, int kmax) {
    const size_t esz = dtype == FT8_AUDIO_I16 ? 2 : 4;
    for (int b0 = 0; b0 < B; b0 += (int)h->y_cycles) {
        const int nb = std::min<int>((int)h->y_cycles, B - b0);
        const char* a = (const char*)d_audio + (size_t)b0 * CYCLE_SAMPLES * esz;
        const int smem = CS_COLS * CS_N1 * (int)sizeof(float2);
        if (dtype == FT8_AUDIO_I16)
            k_cs_cols<int16_t><<<dim3(CS_N2 / CS_COLS, nb), CS_NT, smem, h->stream>>>((const int16_t*)a, h->d_Y, h->d_TC, h->d_W96000T);
        else
            k_cs_cols<float><<<dim3(CS_N2 / CS_COLS, nb), CS_NT, smem, h->stream>>>((const float*)a, h->d_Y, h->d_TC, h->d_W96000T);
        CK(cudaGetLastError());
        k_cs_rows<<<dim3(24, nb), CS_NT, 0, h->stream>>>(h->d_Y, d_spec + (size_t)b0 * stride, stride, kmax, h->d_T256, h->d_W192000);
        CK(cudaGetLastError());
    }
    return FT8_OK;
}

static int persistent_blocks(ft8_handle* h, int per_sm) { return h->n_sm * per_sm; }

// F2/F3 for a work list (list/count on the device, at most cap_slots items) or for items 0..n_direct-1 (list == nullptr).
// fine_mode 0: time scan -> tensor-core frequency scan -> final transform (fine_tc.cuh); 1: the literal 9-transform kernel.
static int launch_fine(ft8_handle* h, const float2* spec, int spec_stride, const int32_t* list, const int32_t* count, int n_direct,
                       const int32_t* cycle_of, const int16_t* f0, const int16_t* h0, FineOut* fo, float* llr, float* sig_grid) {
    if (h->cfg.fine_mode == 1) {
        const int blocks = list ? persistent_blocks(h, 4) : std::min(persistent_blocks(h, 4), n_direct);
        k_fine<<<blocks, FINE_NT, FINE_SMEM_BYTES, h->stream>>>(spec, spec_stride, list, count, n_direct, cycle_of, f0, h0, h->d_TF, fo, llr, sig_grid);
        CK(cudaGetLastError());
        return FT8_OK;
    }
    const size_t need = list ? h->cap_slots : (size_t)n_direct;
    if (need > h->fine_tmp_items) {
        CK(cudaStreamSynchronize(h->stream));
        if (h->d_zwin) CK(cudaFree(h->d_zwin)); if (h->d_tso) CK(cudaFree(h->d_tso)); if (h->d_ff) CK(cudaFree(h->d_ff));
        if (h->d_amb) CK(cudaFree(h->d_amb));
        h->d_zwin = nullptr; h->d_tso = nullptr; h->d_ff = nullptr; h->d_amb = nullptr; h->fine_tmp_items = 0;
        CK(dmalloc(&h->d_zwin, need * FS_WIN)); CK(dmalloc(&h->d_tso, need)); CK(dmalloc(&h->d_ff, need)); CK(dmalloc(&h->d_amb, need + 1));
        h->fine_tmp_items = need;
    }
    const int nbt = list ? persistent_blocks(h, FT_CTAS) : std::min(persistent_blocks(h, FT_CTAS), n_direct);
    const int nb3 = list ? persistent_blocks(h, FF_CTAS) : std::min(persistent_blocks(h, FF_CTAS), n_direct);
    k_fine_tscan<FT_NT, FT_CTAS><<<nbt, FT_NT, FT_SMEM_BYTES, h->stream>>>(spec, spec_stride, list, count, n_direct, cycle_of, f0, h0, h->d_TF, h->d_tso, h->d_zwin);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[12], h->stream));
    const int nb1 = list ? h->n_sm : std::min(h->n_sm, (n_direct + FS_CAND - 1) / FS_CAND);
    CK(cudaMemsetAsync(h->d_amb, 0, sizeof(int32_t), h->stream));
    k_fscan_mma<<<nb1, FS_NT, FS_SMEM_BYTES, h->stream>>>(spec, spec_stride, list, count, n_direct, cycle_of, f0, h0, h->d_tso, h->d_zwin,
                                                         h->d_bmat, h->d_w6400, h->d_ff, h->d_amb + 1, h->d_amb);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[13], h->stream));
    k_fine_final<FT_NT, FF_CTAS><<<nb3, FT_NT, FF_SMEM_BYTES, h->stream>>>(spec, spec_stride, list, count, n_direct, cycle_of, f0, h0, h->d_TF, h->d_tso, h->d_ff,
                                                            fo, llr, sig_grid);
    CK(cudaGetLastError());
    // near-tie candidates of the frequency scan: decided by the literal nine-transform kernel (overwrites their results)
    k_fine<<<persistent_blocks(h, 2), FINE_NT, FINE_SMEM_BYTES, h->stream>>>(spec, spec_stride, h->d_amb + 1, h->d_amb, 0, cycle_of, f0, h0, h->d_TF, fo, llr, sig_grid);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->d_counts + 3, h->d_amb, sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));      // statistics: counts[3]
    return FT8_OK;
}
static int fine_launches(ft8_handle* h) { return h->cfg.fine_mode == 1 ? 1 : 4; }

// copy helpers honouring the mem flag
static int to_device(ft8_handle* h, void* d, const void* src, size_t bytes, int mem) {
    CK(cudaMemcpyAsync(d, src, bytes, mem == FT8_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
    return FT8_OK;
}
static int from_device(ft8_handle* h, void* dst, const void* d, size_t bytes, int mem) {
    CK(cudaMemcpyAsync(dst, d, bytes, mem == FT8_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
    return FT8_OK;
}

#define ENTER(h)                                  \
    if (!(h)) return FT8_E_BADARG;                \
    CK(cudaSetDevice((h)->device));
#define TRY(x) do { int r__ = (x); if (r__ != FT8_OK) return r__; } while (0)

static int stage_audio(ft8_handle* h, const void* audio, int dtype, int B, int mem, const void** d_audio) {
    if (dtype != FT8_AUDIO_I16 && dtype != FT8_AUDIO_F32) return fail(h, FT8_E_BADARG, "audio_dtype must be FT8_AUDIO_I16 or FT8_AUDIO_F32");
    if (mem == FT8_MEM_DEVICE) { *d_audio = audio; return FT8_OK; }
    const size_t bytes = (size_t)B * CYCLE_SAMPLES * (dtype == FT8_AUDIO_I16 ? 2 : 4);
    TRY(ensure_audio(h, bytes));
    TRY(to_device(h, h->d_audio, audio, bytes, mem));
    *d_audio = h->d_audio;
    return FT8_OK;
}

// ------------------------------------------------------------------------------------------ stage ops
extern "C" int ft8_spectrogram(ft8_handle* h, const void* audio, int audio_dtype, int B, float* grid_db, int mem) {
    ENTER(h);
    if (!audio || !grid_db || B <= 0) return fail(h, FT8_E_BADARG, "ft8_spectrogram: bad argument");
    if ((size_t)B > h->cap_cycles) return fail(h, FT8_E_CAPACITY, "ft8_spectrogram: B exceeds cfg.max_cycles");
    const void* da;
    TRY(stage_audio(h, audio, audio_dtype, B, mem, &da));
    float* dg = mem == FT8_MEM_DEVICE ? grid_db : h->d_grid;
    CK(cudaEventRecord(h->ev[10], h->stream));
    TRY(launch_spectrogram(h, da, audio_dtype, B, dg));
    CK(cudaEventRecord(h->ev[11], h->stream));
    if (mem == FT8_MEM_HOST) TRY(from_device(h, grid_db, dg, (size_t)B * GRID_ROWS * GRID_COLS * sizeof(float), mem));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->last_ms[1], h->ev[10], h->ev[11]));
    return FT8_OK;
}

extern "C" int ft8_hop_spectrum(ft8_handle* h, const void* audio_buffer, int audio_dtype, float* row_db, int mem) {
    ENTER(h);
    if (!audio_buffer || !row_db) return fail(h, FT8_E_BADARG, "ft8_hop_spectrum: bad argument");
    if (audio_dtype != FT8_AUDIO_I16 && audio_dtype != FT8_AUDIO_F32) return fail(h, FT8_E_BADARG, "audio_dtype must be FT8_AUDIO_I16 or FT8_AUDIO_F32");
    const size_t esz = audio_dtype == FT8_AUDIO_I16 ? 2 : 4;
    const void* da = audio_buffer;
    if (mem == FT8_MEM_HOST) {
        // only the last 3840 samples are read (receiver.py:289): stage those 15 KB at their place in a cycle-sized buffer,
        // not the whole 720 KB ring, every 40 ms hop
        TRY(ensure_audio(h, (size_t)CYCLE_SAMPLES * esz));
        const size_t off = (size_t)(CYCLE_SAMPLES - NFFT_S) * esz;
        TRY(to_device(h, (char*)h->d_audio + off, (const char*)audio_buffer + off, (size_t)NFFT_S * esz, mem));
        da = h->d_audio;
    }
    float* dr = row_db;
    if (mem == FT8_MEM_HOST) { TRY(ensure_arena(h, GRID_COLS * sizeof(float))); dr = (float*)h->arena; }
    // the window over the last 3840 samples of a 180000-sample buffer is row 375 of that buffer's waterfall
    TRY(launch_spectrogram(h, da, audio_dtype, 1, dr, 375, 375, 1, 375, 0));
    if (mem == FT8_MEM_HOST) TRY(from_device(h, row_db, dr, GRID_COLS * sizeof(float), mem));
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

extern "C" int ft8_sync(ft8_handle* h, const float* grid_db, int grid_rows, int B, int odd_even, int16_t* cand_f0,
                        int16_t* cand_h0, float* cand_score, int32_t* n_cand, float* payload_db, int mem) {
    ENTER(h);
    if (!grid_db || !cand_f0 || !cand_h0 || !cand_score || !n_cand || B <= 0) return fail(h, FT8_E_BADARG, "ft8_sync: bad argument");
    if (grid_rows != GRID_ROWS && grid_rows != LIVE_ROWS) return fail(h, FT8_E_BADARG, "ft8_sync: grid_rows must be 376 or 750");
    if ((size_t)B > h->cap_cycles) return fail(h, FT8_E_CAPACITY, "ft8_sync: B exceeds cfg.max_cycles");
    const size_t K = h->cfg.max_cands, N = (size_t)B * K;
    const float* dg = grid_db;
    if (mem == FT8_MEM_HOST) {
        const size_t bytes = (size_t)B * grid_rows * GRID_COLS * sizeof(float);
        if (grid_rows == GRID_ROWS) { TRY(to_device(h, h->d_grid, grid_db, bytes, mem)); dg = h->d_grid; }
        else { TRY(ensure_arena(h, bytes)); TRY(to_device(h, h->arena, grid_db, bytes, mem)); dg = (const float*)h->arena; }
    }
    CK(cudaMemsetAsync(h->d_f0, 0, N * 2, h->stream));
    CK(cudaMemsetAsync(h->d_h0, 0, N * 2, h->stream));
    CK(cudaMemsetAsync(h->d_score, 0, N * 4, h->stream));
    CK(cudaEventRecord(h->ev[10], h->stream));
    TRY(launch_sync(h, dg, grid_rows, B, odd_even));
    CK(cudaEventRecord(h->ev[11], h->stream));
    float* d_pay = nullptr;
    if (payload_db) {
        if (mem == FT8_MEM_DEVICE) d_pay = payload_db;
        else {
            // the arena may hold the live grid: place the payload scratch after it
            const size_t gbytes = (grid_rows == GRID_ROWS) ? 0 : (size_t)B * grid_rows * GRID_COLS * sizeof(float);
            if (gbytes) return fail(h, FT8_E_BADARG, "ft8_sync: payload_db with a host 750-row grid is not supported; pass grid_rows=376 or device memory");
            TRY(ensure_arena(h, N * 58 * 8 * sizeof(float)));
            d_pay = (float*)h->arena;
        }
        CK(cudaMemsetAsync(d_pay, 0, N * 58 * 8 * sizeof(float), h->stream));
        CK(cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats), h->stream));
        CK(cudaMemsetAsync(h->d_counts, 0, 8 * sizeof(int32_t), h->stream));
        k_pass0<<<persistent_blocks(h, 8), WARPS_PER_CTA * 32, sizeof(PassSmem), h->stream>>>(
            cand_state(h), (int)N, dg, grid_rows, odd_even ? 375 : 0, h->cfg.llr_sd_min, d_pay, 1, h->d_list_fine, h->d_counts, h->d_counts + 6, h->d_stats);
        CK(cudaGetLastError());
    }
    TRY(from_device(h, cand_f0, h->d_f0, N * 2, mem));
    TRY(from_device(h, cand_h0, h->d_h0, N * 2, mem));
    TRY(from_device(h, cand_score, h->d_score, N * 4, mem));
    TRY(from_device(h, n_cand, h->d_ncand, (size_t)B * 4, mem));
    if (payload_db && mem == FT8_MEM_HOST) TRY(from_device(h, payload_db, d_pay, N * 58 * 8 * sizeof(float), mem));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->last_ms[2], h->ev[10], h->ev[11]));
    return FT8_OK;
}

// carve typed sub-buffers out of the arena
struct Carver {
    char* base; size_t off = 0;
    template <typename T> T* take(size_t n) { off = (off + 255) & ~(size_t)255; T* p = (T*)(base + off); off += n * sizeof(T); return p; }
};

extern "C" int ft8_llr(ft8_handle* h, const float* payload_db, int N, float* llr, float* sd, int32_t* snr, int mem) {
    ENTER(h);
    if (!payload_db || !llr || !sd || !snr || N <= 0) return fail(h, FT8_E_BADARG, "ft8_llr: bad argument");
    const float* dp = payload_db; float* dl = llr; float* ds = sd; int32_t* dn = snr;
    if (mem == FT8_MEM_HOST) {
        TRY(ensure_arena(h, (size_t)N * (464 + 174 + 2) * 4 + 4096));
        Carver c{(char*)h->arena};
        float* p = c.take<float>((size_t)N * 464); dl = c.take<float>((size_t)N * 174); ds = c.take<float>(N); dn = c.take<int32_t>(N);
        TRY(to_device(h, p, payload_db, (size_t)N * 464 * 4, mem));
        dp = p;
    }
    k_llr_batch<<<std::min(persistent_blocks(h, 8), (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, h->stream>>>(dp, N, dl, ds, dn);
    CK(cudaGetLastError());
    if (mem == FT8_MEM_HOST) {
        TRY(from_device(h, llr, dl, (size_t)N * 174 * 4, mem));
        TRY(from_device(h, sd, ds, (size_t)N * 4, mem));
        TRY(from_device(h, snr, dn, (size_t)N * 4, mem));
    }
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

extern "C" int ft8_cycle_spectrum(ft8_handle* h, const void* audio, int audio_dtype, int B, float* spec, int mem) {
    ENTER(h);
    if (!audio || !spec || B <= 0) return fail(h, FT8_E_BADARG, "ft8_cycle_spectrum: bad argument");
    if ((size_t)B > h->cap_cycles) return fail(h, FT8_E_CAPACITY, "ft8_cycle_spectrum: B exceeds cfg.max_cycles");
    const void* da;
    TRY(stage_audio(h, audio, audio_dtype, B, mem, &da));
    float2* ds = (float2*)spec;
    if (mem == FT8_MEM_HOST) { TRY(ensure_arena(h, (size_t)B * FT8_SPEC_BINS * sizeof(float2))); ds = (float2*)h->arena; }
    TRY(launch_cycle_spectrum(h, da, audio_dtype, B, ds, FT8_SPEC_BINS, FT8_SPEC_BINS - 1));
    if (mem == FT8_MEM_HOST) TRY(from_device(h, spec, ds, (size_t)B * FT8_SPEC_BINS * sizeof(float2), mem));
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

// ft8_fine helpers: range check of caller-supplied candidates on the device (works for host and device callers alike)
// and the split of FineOut into the ABI's plain arrays.
__global__ void k_fine_validate(const int32_t* __restrict__ cycle_of, const int16_t* __restrict__ f0, const int16_t* __restrict__ h0,
                                int N, int B, int32_t* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && (h0[i] < -100 || h0[i] > 200 || f0[i] < 4 || f0[i] > 1880 || cycle_of[i] < 0 || cycle_of[i] >= B)) atomicAdd(bad, 1);
}
__global__ void k_fine_unpack(const FineOut* __restrict__ fo, int N, int32_t* __restrict__ tt, int32_t* __restrict__ ff,
                              int32_t* __restrict__ ns, float* __restrict__ sd, int32_t* __restrict__ snr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) { const FineOut o = fo[i]; tt[i] = o.tt; ff[i] = o.ff; ns[i] = o.nsync; sd[i] = o.sd; snr[i] = o.snr; }
}

extern "C" int ft8_fine(ft8_handle* h, const float* spec, int B, const int32_t* cycle_of, const int16_t* f0_idx,
                        const int16_t* h0_idx, int N, int32_t* ttweak, int32_t* ftweak, int32_t* nsync, float* signal_grid,
                        float* llr, float* sd, int32_t* snr, int mem) {
    ENTER(h);
    if (!spec || !cycle_of || !f0_idx || !h0_idx || !ttweak || !ftweak || !nsync || !llr || !sd || !snr || N <= 0 || B <= 0)
        return fail(h, FT8_E_BADARG, "ft8_fine: bad argument");
    const size_t specn = (size_t)B * FT8_SPEC_BINS;
    size_t need = specn * 8 + (size_t)N * (4 + 2 + 2 + sizeof(FineOut) + 174 * 4 + 632 * 4 + 5 * 4) + 16 * 256;
    TRY(ensure_arena(h, need));
    Carver c{(char*)h->arena};
    float2* dspec = c.take<float2>(specn);
    int32_t* dco = c.take<int32_t>(N); int16_t* df0 = c.take<int16_t>(N); int16_t* dh0 = c.take<int16_t>(N);
    FineOut* dfo = c.take<FineOut>(N); float* dllr = c.take<float>((size_t)N * 174); float* dsg = c.take<float>((size_t)N * 632);
    int32_t* dtt = c.take<int32_t>(N); int32_t* dff = c.take<int32_t>(N); int32_t* dns = c.take<int32_t>(N);
    float* dsd = c.take<float>(N); int32_t* dsn = c.take<int32_t>(N); int32_t* dbad = c.take<int32_t>(1);
    const float2* sp = (const float2*)spec;
    if (mem == FT8_MEM_HOST) { TRY(to_device(h, dspec, spec, specn * 8, mem)); sp = dspec; }
    TRY(to_device(h, dco, cycle_of, (size_t)N * 4, mem));
    TRY(to_device(h, df0, f0_idx, (size_t)N * 2, mem));
    TRY(to_device(h, dh0, h0_idx, (size_t)N * 2, mem));
    // candidates are range-checked on the device before anything gathers with them (host AND device callers)
    CK(cudaMemsetAsync(dbad, 0, 4, h->stream));
    k_fine_validate<<<(N + 255) / 256, 256, 0, h->stream>>>(dco, df0, dh0, N, B, dbad);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_counts + 3, dbad, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_counts[3] != 0)
        return fail(h, FT8_E_BADARG, "ft8_fine: candidate outside the supported range (-100 <= h0 <= 200, 4 <= f0 <= 1880, 0 <= cycle_of < B)");
    float* out_llr = mem == FT8_MEM_HOST ? dllr : llr;
    float* out_sg = signal_grid ? (mem == FT8_MEM_HOST ? dsg : signal_grid) : nullptr;
    TRY(launch_fine(h, sp, FT8_SPEC_BINS, nullptr, nullptr, N, dco, df0, dh0, dfo, out_llr, out_sg));
    const bool host = mem == FT8_MEM_HOST;
    k_fine_unpack<<<(N + 255) / 256, 256, 0, h->stream>>>(dfo, N, host ? dtt : ttweak, host ? dff : ftweak, host ? dns : nsync,
                                                          host ? dsd : sd, host ? dsn : snr);
    CK(cudaGetLastError());
    if (host) {
        TRY(from_device(h, ttweak, dtt, (size_t)N * 4, mem)); TRY(from_device(h, ftweak, dff, (size_t)N * 4, mem));
        TRY(from_device(h, nsync, dns, (size_t)N * 4, mem)); TRY(from_device(h, sd, dsd, (size_t)N * 4, mem));
        TRY(from_device(h, snr, dsn, (size_t)N * 4, mem));
        TRY(from_device(h, llr, dllr, (size_t)N * 174 * 4, mem));
        if (signal_grid) TRY(from_device(h, signal_grid, dsg, (size_t)N * 632 * 4, mem));
    }
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

extern "C" int ft8_ldpc(ft8_handle* h, float* llr, int N, int max_ncheck0, int max_iters, int32_t* status, int32_t* nits,
                        uint32_t* bits91, int mem) {
    ENTER(h);
    if (!llr || !status || !nits || !bits91 || N <= 0 || max_iters < 0) return fail(h, FT8_E_BADARG, "ft8_ldpc: bad argument");
    float* dl = llr; int32_t* dst = status; int32_t* dn = nits; uint32_t* db = bits91;
    if (mem == FT8_MEM_HOST) {
        TRY(ensure_arena(h, (size_t)N * (174 + 1 + 1 + 3) * 4 + 4096));
        Carver c{(char*)h->arena};
        dl = c.take<float>((size_t)N * 174); dst = c.take<int32_t>(N); dn = c.take<int32_t>(N); db = c.take<uint32_t>((size_t)N * 3);
        TRY(to_device(h, dl, llr, (size_t)N * 174 * 4, mem));
    }
    CK(cudaEventRecord(h->ev[10], h->stream));
    k_ldpc_batch<<<std::min(persistent_blocks(h, 8), (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * 32, sizeof(PassSmem), h->stream>>>(
        dl, N, max_ncheck0, max_iters, dst, dn, db);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[11], h->stream));
    if (mem == FT8_MEM_HOST) {
        TRY(from_device(h, llr, dl, (size_t)N * 174 * 4, mem));
        TRY(from_device(h, status, dst, (size_t)N * 4, mem));
        TRY(from_device(h, nits, dn, (size_t)N * 4, mem));
        TRY(from_device(h, bits91, db, (size_t)N * 12, mem));
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->last_ms[0], h->ev[10], h->ev[11]));
    return FT8_OK;
}

extern "C" int ft8_osd(ft8_handle* h, const float* llr, int N, int singleflips, int doubleflips, int32_t* found,
                       uint32_t* bits91, int mem) {
    ENTER(h);
    if (!llr || !found || !bits91 || N <= 0 || singleflips < 0 || singleflips > OSD_MAX_FLIPS || doubleflips < 0)
        return fail(h, FT8_E_BADARG, "ft8_osd: bad argument (singleflips must be 0..91)");
    const float* dl = llr; int32_t* df = found; uint32_t* db = bits91;
    if (mem == FT8_MEM_HOST) {
        TRY(ensure_arena(h, (size_t)N * (174 + 1 + 3) * 4 + 4096));
        Carver c{(char*)h->arena};
        float* l = c.take<float>((size_t)N * 174); df = c.take<int32_t>(N); db = c.take<uint32_t>((size_t)N * 3);
        TRY(to_device(h, l, llr, (size_t)N * 174 * 4, mem));
        dl = l;
    }
    CK(cudaEventRecord(h->ev[10], h->stream));
    k_osd_batch<<<std::min(persistent_blocks(h, 8), (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA), WARPS_PER_CTA * 32, sizeof(OsdSmem), h->stream>>>(
        dl, N, singleflips, doubleflips, df, db);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[11], h->stream));
    if (mem == FT8_MEM_HOST) {
        TRY(from_device(h, found, df, (size_t)N * 4, mem));
        TRY(from_device(h, bits91, db, (size_t)N * 12, mem));
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->last_ms[0], h->ev[10], h->ev[11]));
    return FT8_OK;
}

extern "C" int ft8_crc14(ft8_handle* h, const uint32_t* bits91, int N, int32_t* flags, int mem) {
    ENTER(h);
    if (!bits91 || !flags || N <= 0) return fail(h, FT8_E_BADARG, "ft8_crc14: bad argument");
    const uint32_t* db = bits91; int32_t* df = flags;
    if (mem == FT8_MEM_HOST) {
        TRY(ensure_arena(h, (size_t)N * 16 + 4096));
        Carver c{(char*)h->arena};
        uint32_t* b = c.take<uint32_t>((size_t)N * 3); df = c.take<int32_t>(N);
        TRY(to_device(h, b, bits91, (size_t)N * 12, mem));
        db = b;
    }
    k_crc_batch<<<std::min(persistent_blocks(h, 8), (N + 127) / 128), 128, 0, h->stream>>>(db, N, df);
    CK(cudaGetLastError());
    if (mem == FT8_MEM_HOST) TRY(from_device(h, flags, df, (size_t)N * 4, mem));
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

// ------------------------------------------------------------------------------------------ whole path
// Record packaging on the device (R1; receiver.py:51-66, 389-398).  One CTA per cycle orders that cycle's decoded
// candidates in the reference's emission order -- pass by pass; inside a pass by llr_sd descending (the fine sd from
// ipass 2 on, all-equal before), ties by candidate rank -- flags later duplicates of a payload (receiver.py:53-55) and
// writes the records contiguously at base[cycle], so the host only copies.
constexpr int REC_MAXC = 928, REC_NT = 256;

__global__ void __launch_bounds__(REC_NT) k_rec_count(CandState cs, int32_t* __restrict__ n_per_cycle) {
    __shared__ int cnt;
    const int cyc = blockIdx.x;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int n = cs.n_cand[cyc];
    int c = 0;
    for (int r = threadIdx.x; r < n; r += REC_NT) c += cs.status[cyc * cs.K + r] == ST_DECODED;
    if (c) atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) n_per_cycle[cyc] = cnt;
}

// exclusive scan of n_per_cycle[B] -> base[B], total -> *total (single CTA)
__global__ void __launch_bounds__(1024) k_rec_scan(const int32_t* __restrict__ n_per_cycle, int B, int32_t* __restrict__ base,
                                                   int32_t* __restrict__ total) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < B; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < B ? n_per_cycle[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const int before = carry + (w ? warp_sum[w - 1] : 0) + x - v;
        if (i < B) base[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(REC_NT)
k_rec_write(CandState cs, const FineOut* __restrict__ fine, const int32_t* __restrict__ base, ft8_record* __restrict__ rec,
            DevStats* __restrict__ stats) {
    __shared__ uint16_t s_rank[REC_MAXC];      // candidate rank (index inside the cycle) of the i-th decoded candidate
    __shared__ uint8_t s_ipass[REC_MAXC];
    __shared__ float s_sd[REC_MAXC];
    __shared__ uint32_t s_w[REC_MAXC][3];
    __shared__ int s_n, s_emit;
    const int cyc = blockIdx.x, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_n = 0; s_emit = 0; }
    __syncthreads();
    const int n = cs.n_cand[cyc];
    // compaction in candidate order (ballot scan per warp-sized chunk keeps it deterministic)
    for (int r0 = 0; r0 < n; r0 += REC_NT) {
        const int r = r0 + threadIdx.x;
        const bool dec = r < n && cs.status[cyc * cs.K + r] == ST_DECODED;
        const uint32_t m = __ballot_sync(0xffffffffu, dec);
        __shared__ int wbase[REC_NT / 32];
        if (lane == 0) wbase[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        int off = s_n;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) off += wbase[w];
        if (dec) {
            const int i = off + __popc(m & ((1u << lane) - 1u));
            const int slot = cyc * cs.K + r;
            s_rank[i] = (uint16_t)r;
            s_ipass[i] = cs.r_ipass[slot];
            s_sd[i] = cs.r_ipass[slot] >= 2 ? fine[slot].sd : 0.0f;
            s_w[i][0] = cs.bits91[3 * slot]; s_w[i][1] = cs.bits91[3 * slot + 1]; s_w[i][2] = cs.bits91[3 * slot + 2] & 0x1FFFu;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < REC_NT / 32; ++w) t += wbase[w]; s_n += t; }
        __syncthreads();
    }
    const int nd = s_n;
    int emitted_here = 0;
    for (int i = threadIdx.x; i < nd; i += REC_NT) {
        const int ip = s_ipass[i];
        const float sd = s_sd[i];
        const int rk = s_rank[i];
        int pos = 0;
        bool dup = false;
        for (int j = 0; j < nd; ++j) {
            const int jp = s_ipass[j];
            bool before;                      // does j precede i in emission order?
            if (jp != ip) before = jp < ip;
            else if (ip >= 2 && s_sd[j] != sd) before = s_sd[j] > sd;
            else before = s_rank[j] < rk;
            if (before) {
                ++pos;
                dup = dup || (s_w[j][0] == s_w[i][0] && s_w[j][1] == s_w[i][1] && s_w[j][2] == s_w[i][2]);
            }
        }
        const int slot = cyc * cs.K + rk;
        ft8_record r;
        memset(&r, 0, sizeof(r));
        r.bits91[0] = cs.bits91[3 * slot]; r.bits91[1] = cs.bits91[3 * slot + 1]; r.bits91[2] = cs.bits91[3 * slot + 2];
        r.cycle = cyc; r.cand = (int16_t)rk; r.f0_idx = cs.f0[slot]; r.h0_idx = cs.h0[slot];
        r.ipass = (uint8_t)ip; r.ap = cs.r_ap[slot]; r.method = cs.r_method[slot]; r.n_its = cs.r_nits[slot];
        r.score = cs.score[slot]; r.grid_sd = cs.grid_sd[slot];
        r.emitted = dup ? 0 : 1;
        emitted_here += dup ? 0 : 1;
        // float(h0/25) and 3.125*f0 like receiver.py:350-351 (python floats; rounded to fp32 for the record)
        double tsec = (double)r.h0_idx / 25.0, fhz = 3.125 * (double)r.f0_idx;
        if (ip >= 2) {
            const FineOut fo = fine[slot];
            r.ttweak = (int8_t)fo.tt; r.ftweak = (int8_t)fo.ff; r.nsync = (uint8_t)fo.nsync; r.fine_sd = fo.sd; r.snr = (int8_t)fo.snr;
            tsec += (double)fo.tt / 200.0; fhz += (double)fo.ff / 16.0;
        } else {
            r.nsync = 100; r.fine_sd = nanf(""); r.snr = cs.grid_snr[slot];
        }
        r.tsec = (float)tsec; r.fHz = (float)fhz;
        rec[base[cyc] + pos] = r;
    }
    if (emitted_here) atomicAdd(&s_emit, emitted_here);
    __syncthreads();
    if (threadIdx.x == 0 && s_emit) atomicAdd(&stats->emitted, (unsigned long long)s_emit);
}

extern "C" int ft8_prefetch_audio(ft8_handle* h, const void* audio_host, int audio_dtype, int B) {
    ENTER(h);
    if (!audio_host || B <= 0 || (audio_dtype != FT8_AUDIO_I16 && audio_dtype != FT8_AUDIO_F32))
        return fail(h, FT8_E_BADARG, "ft8_prefetch_audio: bad argument");
    if ((size_t)B > h->cap_cycles) return fail(h, FT8_E_CAPACITY, "ft8_prefetch_audio: B exceeds cfg.max_cycles");
    const size_t bytes = (size_t)B * CYCLE_SAMPLES * (audio_dtype == FT8_AUDIO_I16 ? 2 : 4);
    if (bytes > h->audio_pf_bytes) {
        if (h->d_audio_pf) { CK(cudaStreamSynchronize(h->copy_stream)); CK(cudaStreamSynchronize(h->stream)); CK(cudaFree(h->d_audio_pf)); }
        h->d_audio_pf = nullptr; h->audio_pf_bytes = 0;
        CK(cudaMalloc(&h->d_audio_pf, bytes));
        h->audio_pf_bytes = bytes;
    }
    // the slot may still be read by kernels of the call that consumed the previous prefetch (same stream order)
    CK(cudaEventRecord(h->pf_ev, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->pf_ev, 0));
    CK(cudaMemcpyAsync(h->d_audio_pf, audio_host, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->pf_ev, h->copy_stream));
    h->pf_host = audio_host; h->pf_B = B; h->pf_dtype = audio_dtype;
    return FT8_OK;
}

// the previous cycle's last 3840 samples per stream, kept for the next live call (D2D strided copy on the handle's stream)
static int save_tails(ft8_handle* h, const void* d_audio, int dtype, int B) {
    const size_t esz = dtype == FT8_AUDIO_I16 ? 2 : 4;
    CK(cudaMemcpy2DAsync(h->d_tail, (size_t)NFFT_S * esz, (const char*)d_audio + (size_t)(CYCLE_SAMPLES - NFFT_S) * esz,
                         (size_t)CYCLE_SAMPLES * esz, (size_t)NFFT_S * esz, (size_t)B, cudaMemcpyDeviceToDevice, h->stream));
    h->tail_dtype = dtype; h->tail_valid = true;
    return FT8_OK;
}

__global__ void k_fill(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static int decode_cycles_core(ft8_handle* h, const void* audio, int audio_dtype, int B, int odd_even, ft8_record* rec,
                              int rec_capacity, int32_t* n_rec, int mem, const void* next_audio_host, bool consume_pf,
                              bool live = false) {
    ENTER(h);
    if (!audio || !rec || !n_rec || B <= 0 || rec_capacity < 0) return fail(h, FT8_E_BADARG, "ft8_decode_cycles: bad argument");
    if ((size_t)B > h->cap_cycles) return fail(h, FT8_E_CAPACITY, "ft8_decode_cycles: B exceeds cfg.max_cycles");
    const int K = h->cfg.max_cands, N = B * K;
    // Isolated mode (default): every cycle is decoded on its own 376-row waterfall, the search runs with
    // cycle_h0 = 0 and odd_even only labels the records on the host side (their_tx_cycle, receiver.py:62).
    // Live mode: stream b keeps the reference's 750-row two-cycle ring (receiver.py:238, 295-306); this call's rows go to the
    // half odd_even selects, rows of the other half still hold the previous cycle, and the first hops' windows reach back into
    // the previous cycle's last samples -- so candidates with h0 < -32 read real rows, exactly as a live Receiver does.
    if (audio_dtype != FT8_AUDIO_I16 && audio_dtype != FT8_AUDIO_F32) return fail(h, FT8_E_BADARG, "audio_dtype must be FT8_AUDIO_I16 or FT8_AUDIO_F32");
    const size_t esz = audio_dtype == FT8_AUDIO_I16 ? 2 : 4;
    const int cycle_h0 = live ? (odd_even ? 375 : 0) : 0;
    const int grows = live ? LIVE_ROWS : GRID_ROWS;
    float* gbase = h->d_grid;
    const void* tail = nullptr;
    if (live) {
        if (!h->d_ring) {
            const size_t n = h->cap_cycles * (size_t)LIVE_ROWS * GRID_COLS;
            CK(dmalloc(&h->d_ring, n));
            k_fill<<<h->n_sm * 8, 256, 0, h->stream>>>(h->d_ring, n, 1.0f);           // np.ones (receiver.py:238)
            CK(cudaGetLastError());
            CK(cudaMalloc(&h->d_tail, h->cap_cycles * (size_t)NFFT_S * 4));
        }
        gbase = h->d_ring;
        if (h->tail_valid && h->tail_dtype == audio_dtype) tail = h->d_tail;
    }
    const void* da = audio;
    // A pending prefetch is consumed only by the streaming entry (consume_pf), whose caller named this very buffer as the
    // next batch and so promises it has not been rewritten since; every other call drops it -- after waiting for its copy,
    // which may still be reading the caller's host buffer.
    const bool prefetched = consume_pf && mem == FT8_MEM_HOST && h->pf_host == audio && h->pf_B == B && h->pf_dtype == audio_dtype;
    if (!prefetched && h->pf_host) CK(cudaStreamSynchronize(h->copy_stream));
    if (prefetched) {
        // this batch was copied by ft8_prefetch_audio while the previous call was computing: swap it in and go on as if
        // the audio were device-resident
        std::swap(h->d_audio, h->d_audio_pf);
        std::swap(h->audio_bytes, h->audio_pf_bytes);
        h->pf_host = nullptr;
        CK(cudaStreamWaitEvent(h->stream, h->pf_ev, 0));
        da = h->d_audio;
        mem = FT8_MEM_DEVICE;
    } else {
        h->pf_host = nullptr;                    // a pending prefetch this call does not consume is dropped, never decoded later
        if (mem == FT8_MEM_HOST) { TRY(ensure_audio(h, (size_t)B * CYCLE_SAMPLES * esz)); da = h->d_audio; }
    }
    // streaming callers name the next batch: its copy is queued now and runs underneath this batch's kernels
    // (after this batch's own copies when it was not prefetched itself, so that they are not queued behind it)
    if (next_audio_host && prefetched) TRY(ft8_prefetch_audio(h, next_audio_host, audio_dtype, B));
    int launches = 0;
    CK(cudaMemsetAsync(h->d_counts, 0, 8 * sizeof(int32_t), h->stream));   // [0] fine list, [1] osd list, [2] records, [4],[5] work cursors
    CK(cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats), h->stream));
    CK(cudaEventRecord(h->ev[0], h->stream));
    if (mem == FT8_MEM_HOST && B > 64) {
        // host audio: copy in chunks on the copy stream; S1, S2 and F1 of a chunk start as soon as it has landed,
        // so the PCIe transfer hides behind the front-end kernels (stage timers then cover the whole front end as "S1")
        // up to 16 equal chunks (measured best for a single handle: 28.6 k cycles/s vs 27.3-28.0 k for growing chunks);
        // callers that stream batch after batch should use ft8_prefetch_audio, which hides the whole transfer
        int cb[16], cn[16];
        const int nchunk = std::min(16, (B + 255) / 256);
        const int per = (B + nchunk - 1) / nchunk;
        for (int c = 0; c < nchunk; ++c) { cb[c] = c * per; cn[c] = std::max(0, std::min(per, B - c * per)); }
        CK(cudaEventRecord(h->chunk_ev[0], h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[0], 0));        // previous users of d_audio on the main stream are done
        for (int c = 0; c < nchunk; ++c) {
            if (cn[c] <= 0) continue;
            const size_t off = (size_t)cb[c] * CYCLE_SAMPLES * esz;
            CK(cudaMemcpyAsync((char*)h->d_audio + off, (const char*)audio + off, (size_t)cn[c] * CYCLE_SAMPLES * esz, cudaMemcpyHostToDevice, h->copy_stream));
            CK(cudaEventRecord(h->chunk_ev[c], h->copy_stream));
        }
        for (int c = 0; c < nchunk; ++c) {
            const int b0 = cb[c], nb = cn[c];
            if (nb <= 0) continue;
            const char* a = (const char*)h->d_audio + (size_t)b0 * CYCLE_SAMPLES * esz;
            CK(cudaStreamWaitEvent(h->stream, h->chunk_ev[c], 0));
            float* gc = gbase + (size_t)b0 * grows * GRID_COLS;
            const void* tc = tail ? (const char*)tail + (size_t)b0 * NFFT_S * esz : nullptr;
            if (live) TRY(launch_spectrogram(h, a, audio_dtype, nb, gc, 1, 375, LIVE_ROWS, -cycle_h0, 0, tc, 1));
            else TRY(launch_spectrogram(h, a, audio_dtype, nb, gc));
            ++launches;
            TRY(launch_sync(h, gc, grows, nb, live ? odd_even : 0, b0)); launches += 2;
            TRY(launch_cycle_spectrum(h, a, audio_dtype, nb, h->d_spec + (size_t)b0 * FINE_SPEC_STRIDE, FINE_SPEC_STRIDE, FINE_SPEC_STRIDE - 1));
            launches += 2 * ((nb + (int)h->y_cycles - 1) / (int)h->y_cycles);
        }
        CK(cudaEventRecord(h->ev[1], h->stream));
        CK(cudaEventRecord(h->ev[2], h->stream));
        CK(cudaEventRecord(h->ev[3], h->stream));
    } else {
        if (mem == FT8_MEM_HOST) TRY(to_device(h, h->d_audio, audio, (size_t)B * CYCLE_SAMPLES * esz, mem));
        // S1
        if (live) TRY(launch_spectrogram(h, da, audio_dtype, B, gbase, 1, 375, LIVE_ROWS, -cycle_h0, 0, tail, 1));
        else TRY(launch_spectrogram(h, da, audio_dtype, B, gbase));
        ++launches;
        CK(cudaEventRecord(h->ev[1], h->stream));
        // S2
        TRY(launch_sync(h, gbase, grows, B, live ? odd_even : 0)); launches += 2;
        CK(cudaEventRecord(h->ev[2], h->stream));
        // F1 (independent of S1/S2; same stream)
        TRY(launch_cycle_spectrum(h, da, audio_dtype, B, h->d_spec, FINE_SPEC_STRIDE, FINE_SPEC_STRIDE - 1));
        launches += 2 * ((B + (int)h->y_cycles - 1) / (int)h->y_cycles);
        CK(cudaEventRecord(h->ev[3], h->stream));
    }
    if (next_audio_host && !prefetched) TRY(ft8_prefetch_audio(h, next_audio_host, audio_dtype, B));
    if (live) TRY(save_tails(h, mem == FT8_MEM_HOST ? h->d_audio : da, audio_dtype, B));      // after S1 / F1 have read this audio
    CandState cs = cand_state(h);
    // ipass 0
    k_pass0<<<persistent_blocks(h, 8), WARPS_PER_CTA * 32, sizeof(PassSmem), h->stream>>>(
        cs, N, gbase, grows, cycle_h0, h->cfg.llr_sd_min, nullptr, 0, h->d_list_fine, h->d_counts + 0, h->d_counts + 6, h->d_stats);
    CK(cudaGetLastError()); ++launches;
    CK(cudaEventRecord(h->ev[4], h->stream));
    // ipass 1
    TRY(launch_fine(h, h->d_spec, FINE_SPEC_STRIDE, h->d_list_fine, h->d_counts + 0, 0, h->d_cycle_of, h->d_f0, h->d_h0, h->d_fine,
                    h->d_llr_fine, nullptr));
    launches += fine_launches(h);
    CK(cudaEventRecord(h->ev[5], h->stream));
    // ipass 2-4
    k_pass234<<<persistent_blocks(h, 8), WARPS_PER_CTA * 32, sizeof(PassSmem), h->stream>>>(
        cs, h->d_list_fine, h->d_counts + 0, h->cfg.llr_sd_min, h->d_list_osd, h->d_counts + 1, h->d_counts + 4, h->d_stats);
    CK(cudaGetLastError()); ++launches;
    CK(cudaEventRecord(h->ev[6], h->stream));
    // ipass 5-6
    k_osd_items<<<persistent_blocks(h, 8), WARPS_PER_CTA * 32, sizeof(OsdSmem), h->stream>>>(
        cs, h->d_list_osd, h->d_counts + 1, h->cfg.osd_singleflips, h->cfg.osd_doubleflips, h->d_counts + 5, h->d_stats);
    CK(cudaGetLastError()); ++launches;
    k_osd_resolve<<<persistent_blocks(h, 2), 128, 0, h->stream>>>(cs, h->d_list_osd, h->d_counts + 1, h->d_stats);
    CK(cudaGetLastError()); ++launches;
    CK(cudaEventRecord(h->ev[7], h->stream));
    k_rec_count<<<B, REC_NT, 0, h->stream>>>(cs, h->d_rec_n);
    CK(cudaGetLastError()); ++launches;
    k_rec_scan<<<1, 1024, 0, h->stream>>>(h->d_rec_n, B, h->d_rec_base, h->d_counts + 2);
    CK(cudaGetLastError()); ++launches;
    k_rec_write<<<B, REC_NT, 0, h->stream>>>(cs, h->d_fine, h->d_rec_base, h->d_rec, h->d_stats);
    CK(cudaGetLastError()); ++launches;
    CK(cudaEventRecord(h->ev[8], h->stream));
    CK(cudaMemcpyAsync(h->h_counts, h->d_counts, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(n_rec, h->d_rec_n, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const int nrec = h->h_counts[2];
    const int ncopy = std::min(nrec, rec_capacity);
    if (ncopy > 0) {
        CK(cudaMemcpyAsync(rec, h->d_rec, (size_t)ncopy * sizeof(ft8_record), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const bool overflow = nrec > rec_capacity;
    if (overflow) {                       // records are cycle-major: trim the per-cycle counts to what was copied
        int left = rec_capacity;
        for (int b = 0; b < B; ++b) { const int t = std::min(n_rec[b], left); n_rec[b] = t; left -= t; }
    }
    const int64_t emitted = (int64_t)h->h_stats->emitted;
    CK(cudaEventElapsedTime(&h->last_ms[0], h->ev[0], h->ev[8]));
    for (int i = 1; i <= 8; ++i) CK(cudaEventElapsedTime(&h->last_ms[i], h->ev[i - 1], h->ev[i]));
    h->last_ms[9] = h->last_ms[10] = h->last_ms[11] = 0.0f;
    if (h->cfg.fine_mode == 0) {              // the three kernels of the fine stage: time scan, tensor-core frequency scan, final
        CK(cudaEventElapsedTime(&h->last_ms[9], h->ev[4], h->ev[12]));
        CK(cudaEventElapsedTime(&h->last_ms[10], h->ev[12], h->ev[13]));
        CK(cudaEventElapsedTime(&h->last_ms[11], h->ev[13], h->ev[5]));
    }
    ft8_stats& s = h->stats;
    memset(&s, 0, sizeof(s));
    s.cycles = B; s.candidates = (int64_t)h->h_stats->candidates; s.stopped_sd = (int64_t)h->h_stats->stopped_sd;
    s.fine_evals = (int64_t)h->h_stats->fine_evals; s.fine_pass = (int64_t)h->h_stats->fine_pass;
    s.ldpc_calls = (int64_t)h->h_stats->ldpc_calls; s.ldpc_iters = (int64_t)h->h_stats->ldpc_iters;
    s.osd_calls = (int64_t)h->h_stats->osd_calls; s.decoded = nrec; s.emitted = emitted; s.kernel_launches = launches;
    s.fine_rechecked = h->cfg.fine_mode == 0 ? h->h_counts[3] : 0;
    if (overflow) return fail(h, FT8_E_CAPACITY, "ft8_decode_cycles: rec_capacity too small (records truncated)");
    return FT8_OK;
}

extern "C" int ft8_decode_cycles(ft8_handle* h, const void* audio, int audio_dtype, int B, int odd_even, ft8_record* rec,
                                 int rec_capacity, int32_t* n_rec, int mem) {
    return decode_cycles_core(h, audio, audio_dtype, B, odd_even, rec, rec_capacity, n_rec, mem, nullptr, false);
}

extern "C" int ft8_decode_cycles_live(ft8_handle* h, const void* audio, int audio_dtype, int B, int odd_even, ft8_record* rec,
                                      int rec_capacity, int32_t* n_rec, int mem) {
    if (odd_even != 0 && odd_even != 1) return fail(h, FT8_E_BADARG, "ft8_decode_cycles_live: odd_even must be 0 or 1");
    return decode_cycles_core(h, audio, audio_dtype, B, odd_even, rec, rec_capacity, n_rec, mem, nullptr, false, true);
}

extern "C" int ft8_live_reset(ft8_handle* h) {
    ENTER(h);
    if (h->d_ring) {
        k_fill<<<h->n_sm * 8, 256, 0, h->stream>>>(h->d_ring, h->cap_cycles * (size_t)LIVE_ROWS * GRID_COLS, 1.0f);
        CK(cudaGetLastError());
    }
    h->tail_valid = false;
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

extern "C" int ft8_decode_cycles_stream(ft8_handle* h, const void* audio, int audio_dtype, int B, int odd_even, ft8_record* rec,
                                        int rec_capacity, int32_t* n_rec, const void* next_audio_host) {
    return decode_cycles_core(h, audio, audio_dtype, B, odd_even, rec, rec_capacity, n_rec, FT8_MEM_HOST, next_audio_host, true);
}

// ------------------------------------------------------------------------------------------ generator + test hook
extern "C" int ft8_synth_cycles(ft8_handle* h, const uint8_t* symbols, const float* f_hz, const float* dt_s, const float* amp,
                                int B, int n_sig, float noise_sigma, uint64_t seed, int16_t* audio, int mem) {
    ENTER(h);
    if (!symbols || !f_hz || !dt_s || !amp || !audio || B <= 0 || n_sig < 0 || n_sig > SYNTH_MAX_SIG)
        return fail(h, FT8_E_BADARG, "ft8_synth_cycles: bad argument (n_sig <= 128)");
    const size_t ns = (size_t)B * n_sig;
    TRY(ensure_arena(h, ns * (79 + 12) + (mem == FT8_MEM_HOST ? (size_t)B * CYCLE_SAMPLES * 2 : 0) + 4096));
    Carver c{(char*)h->arena};
    uint8_t* dsym = c.take<uint8_t>(ns * 79); float* df = c.take<float>(ns); float* dd = c.take<float>(ns); float* dam = c.take<float>(ns);
    int16_t* da = mem == FT8_MEM_HOST ? c.take<int16_t>((size_t)B * CYCLE_SAMPLES) : audio;
    // parameters are small and always come from the host
    CK(cudaMemcpyAsync(dsym, symbols, ns * 79, cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(df, f_hz, ns * 4, cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(dd, dt_s, ns * 4, cudaMemcpyDefault, h->stream));
    CK(cudaMemcpyAsync(dam, amp, ns * 4, cudaMemcpyDefault, h->stream));
    k_synth<<<dim3((CYCLE_SAMPLES + SYNTH_TILE - 1) / SYNTH_TILE, B), SYNTH_NT, 0, h->stream>>>(dsym, df, dd, dam, n_sig, noise_sigma, seed, h->d_pulse, da);
    CK(cudaGetLastError());
    if (mem == FT8_MEM_HOST) TRY(from_device(h, audio, da, (size_t)B * CYCLE_SAMPLES * 2, mem));
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

template <int N, int NT, bool INV>
__global__ void k_debug_fft(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ W) {
    extern __shared__ float2 dbg_smem[];
    const float2* x = in + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += NT) dbg_smem[i] = x[i];
    __syncthreads();
    if (N == 1920) fft1920<NT, INV>(dbg_smem, threadIdx.x, W);
    if (N == 3200) fft3200<NT, INV>(dbg_smem, threadIdx.x, W);
    if (N == 375) fft375<NT, INV>(dbg_smem, threadIdx.x, W);
    if (N == 256) fft256<NT, INV>(dbg_smem, threadIdx.x, W);
    if (N == 32) fft32<NT, INV>(dbg_smem, threadIdx.x, W);
    for (int i = threadIdx.x; i < N; i += NT) out[(size_t)blockIdx.x * N + i] = dbg_smem[i];
}

extern "C" int ft8_debug_fft(ft8_handle* h, int n, int inverse, const float* in, float* out, int batch) {
    ENTER(h);
    if (!in || !out || batch <= 0) return fail(h, FT8_E_BADARG, "ft8_debug_fft: bad argument");
    const size_t bytes = (size_t)batch * n * sizeof(float2);
    TRY(ensure_arena(h, 2 * bytes + 512));
    Carver c{(char*)h->arena};
    float2* di = c.take<float2>((size_t)batch * n); float2* dout = c.take<float2>((size_t)batch * n);
    TRY(to_device(h, di, in, bytes, FT8_MEM_HOST));
#define DBG(NN, NT, WT)                                                                                                    \
    if (n == NN) {                                                                                                         \
        if (inverse) k_debug_fft<NN, NT, true><<<batch, NT, NN * sizeof(float2), h->stream>>>(di, dout, WT);               \
        else k_debug_fft<NN, NT, false><<<batch, NT, NN * sizeof(float2), h->stream>>>(di, dout, WT);                      \
    }
    DBG(1920, 128, h->d_W1920) else DBG(3200, 256, h->d_W3200) else DBG(375, 128, h->d_W375) else DBG(256, 64, h->d_W256)
    else DBG(32, 32, h->d_W32) else return fail(h, FT8_E_BADARG, "ft8_debug_fft: n must be 32, 256, 375, 1920 or 3200");
#undef DBG
    CK(cudaGetLastError());
    TRY(from_device(h, out, dout, bytes, FT8_MEM_HOST));
    CK(cudaStreamSynchronize(h->stream));
    return FT8_OK;
}

#ifdef FS_PROFILE
extern "C" int ft8_debug_fs_prof(unsigned long long* out16, int reset) {
    cudaMemcpyFromSymbol(out16, ft8::g_fs_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ft8::g_fs_prof, z, sizeof(z)); }
    return 0;
}
#endif

#ifdef FINE_PROFILE
// debug build only: read (and optionally reset) the k_fine phase counters
extern "C" int ft8_debug_fine_prof(unsigned long long* out16, int reset) {
    cudaMemcpyFromSymbol(out16, ft8::g_fine_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(ft8::g_fine_prof, z, sizeof(z)); }
    return 0;
}
#endif
