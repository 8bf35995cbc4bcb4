// itx.cuh -- AV1 inverse transforms (DCT4..64, ADST4/8/16, identity4..32, WHT4), bit-exact with
// the reference's TransformBlock::inverseTransform (decoder/TransformBlock.cpp:1751-2253).
//
// The DCT and the 8 / 16-point ADST are normative flow graphs (spec 7.13.2.3, .7, .8): they are
// not written out here but GENERATED (tools/gen_itx.py -> itx_gen.h) as straight-line code, one
// function per transform and per number of leading non-zero inputs, with every multiply and add
// that a zero input would feed removed and common sub-expressions shared.  The N-point vector
// lives in registers; row / column passes pick the variant from the op's nz_cols / nz_rows.
// ADST4, the identity transforms and WHT4 are closed formulas and stay below.
#pragma once
#include "dev.h"
#include "av1_tables.h"
#include "itx_gen.h"

namespace itx {

// ADST4 (spec 7.13.2.6; reference iAdst4(), TransformBlock.cpp:1991)
AV1B_DEV void iadst4(int* T)
{
    const int S1 = 1321, S2 = 2482, S3 = 3344, S4 = 3803;
    int s0 = S1 * T[0], s1 = S2 * T[0], s2 = S3 * T[1], s3 = S4 * T[2];
    int s4 = S1 * T[2], s5 = S2 * T[3], s6 = S4 * T[3];
    int b7 = T[0] - T[2] + T[3];
    s0 = s0 + s3 + s5;
    s1 = s1 - s4 - s6;
    s3 = s2;
    s2 = S3 * b7;
    T[0] = (s0 + s3 + 2048) >> 12;
    T[1] = (s1 + s3 + 2048) >> 12;
    T[2] = (s2 + 2048) >> 12;
    T[3] = (s0 + s1 - s3 + 2048) >> 12;
}

// Identity (spec 7.13.2.15; reference iIdentity(), TransformBlock.cpp:2127)
template <int n> AV1B_DEV void iidentity(int* T)
{
    constexpr int N = 1 << n;
    AV1B_UNROLL
    for (int i = 0; i < N; i++) {
        if (n == 2) T[i] = (T[i] * 5793 + 2048) >> 12;
        else if (n == 3) T[i] = T[i] * 2;
        else if (n == 4) T[i] = (T[i] * 11586 + 2048) >> 12;
        else if (n == 5) T[i] = T[i] * 4;
    }
}

// WHT4 (spec 7.13.2.10; reference inverseWalshHadamardTransform(), TransformBlock.cpp:2149)
AV1B_DEV void iwht4(int* T, int shift)
{
    int a = T[0] >> shift, c = T[1] >> shift, d = T[2] >> shift, b = T[3] >> shift;
    a += c;
    d -= b;
    int e = (a - d) >> 1;
    b = e - b;
    c = e - c;
    a -= b;
    d += c;
    T[0] = a;
    T[1] = b;
    T[2] = c;
    T[3] = d;
}

enum { K_DCT = 0, K_ADST = 1, K_IDT = 2, K_WHT = 3 };

// 1-D kernel class of the row (horizontal) and column (vertical) pass for each TX_TYPE
// (the two switch statements at TransformBlock.cpp:2194-2214 and 2228-2248).
AV1B_DEV int row_kind(int tx_type)
{
    // DCT: DCT_DCT ADST_DCT FLIPADST_DCT H_DCT ; IDT: IDTX V_DCT V_ADST V_FLIPADST ; else ADST
    const unsigned dct = (1u << 0) | (1u << 1) | (1u << 4) | (1u << 11);
    const unsigned idt = (1u << 9) | (1u << 10) | (1u << 12) | (1u << 14);
    return ((dct >> tx_type) & 1) ? K_DCT : (((idt >> tx_type) & 1) ? K_IDT : K_ADST);
}
AV1B_DEV int col_kind(int tx_type)
{
    // DCT: DCT_DCT DCT_ADST DCT_FLIPADST V_DCT ; IDT: IDTX H_DCT H_ADST H_FLIPADST ; else ADST
    const unsigned dct = (1u << 0) | (1u << 2) | (1u << 5) | (1u << 10);
    const unsigned idt = (1u << 9) | (1u << 11) | (1u << 13) | (1u << 15);
    return ((dct >> tx_type) & 1) ? K_DCT : (((idt >> tx_type) & 1) ? K_IDT : K_ADST);
}
AV1B_DEV bool flip_ud(int tx_type) { return ((1u << tx_type) & ((1u << 4) | (1u << 8) | (1u << 14) | (1u << 6))) != 0; }
AV1B_DEV bool flip_lr(int tx_type) { return ((1u << tx_type) & ((1u << 5) | (1u << 7) | (1u << 15) | (1u << 6))) != 0; }

// nz: number of leading entries of T that may be non-zero
template <int n> AV1B_DEV void run1d(int* T, int kind, int r, int wht_shift, int nz)
{
    if (kind == K_DCT) {
        if (n == 2) idct4(T, r, nz);
        else if (n == 3) idct8(T, r, nz);
        else if (n == 4) idct16(T, r, nz);
        else if (n == 5) idct32(T, r, nz);
        else idct64(T, r, nz);
    } else if (kind == K_ADST) {
        if (n == 2) iadst4(T);
        else if (n == 3) iadst8(T, r, nz);
        else if (n == 4) iadst16(T, r, nz);
    } else if (kind == K_IDT) {
        if (n <= 5) iidentity<(n <= 5 ? n : 5)>(T);
    } else {
        if (n == 2) iwht4(T, wht_shift);
    }
}

// Row pass of one coefficient row: coef (int16, tw stored entries of which the first tw_nz may be
// non-zero) -> tmp row (int16, N entries).
template <int n>
AV1B_DEV void row_pass(const int16_t* coef_row, int tw, int tw_nz, int16_t* tmp_row, int kind, bool rect, int row_shift)
{
    constexpr int N = 1 << n;
    int T[N];
    AV1B_UNROLL
    for (int j = 0; j < N; j++) T[j] = (j < tw_nz) ? (int)coef_row[j] : 0;
    if (rect) {
        AV1B_UNROLL
        for (int j = 0; j < N; j++) T[j] = (T[j] * 2896 + 2048) >> 12;
    }
    run1d<n>(T, kind, 16, 2, tw_nz);
    AV1B_UNROLL
    for (int j = 0; j < N; j++) tmp_row[j] = (int16_t)clip3(-32768, 32767, round2(T[j], row_shift));
}

// Column pass of column j: tmp (int16, stride tmp_stride, nz_rows valid rows) -> residual.
template <int n>
AV1B_DEV void col_pass(const int16_t* tmp_col, int tmp_stride, int nz_rows, int16_t* out_col, int out_stride,
    bool fud, int kind, int col_shift)
{
    constexpr int N = 1 << n;
    int T[N];
    AV1B_UNROLL
    for (int i = 0; i < N; i++) T[i] = (i < nz_rows) ? (int)tmp_col[i * tmp_stride] : 0;
    run1d<n>(T, kind, 16, 0, nz_rows);
    AV1B_UNROLL
    for (int i = 0; i < N; i++) {
        int v = clip3(-32768, 32767, round2(T[i], col_shift));
        out_col[(fud ? (N - 1 - i) : i) * out_stride] = (int16_t)v;
    }
}

// ---- 32- and 64-point DCT shared by two lanes (of different warps): lane `half` 0 runs the DCT of
// half the size on the even-indexed inputs, lane 1 the odd half of the flow graph on the odd-indexed
// ones; the caller exchanges the results through shared memory and applies the last Hadamard stage.
// Each lane holds N/2 values: half the registers and half the latency of the one-lane transform.
// X: where this lane's N/2 results go (ints, before the last stage).  nz: leading inputs of the
// N-point transform that may be non-zero (<= 32).
template <int n> AV1B_DEV void dct_half(int* T, int half, int cnt)
{
    if (half == 0) {
        if (n == 6) {
            const int hi = 32767, lo = -32768;
            if (cnt <= 1) idct32_k1(T, lo, hi);
            else if (cnt <= 8) idct32_k8(T, lo, hi);
            else idct32_k16(T, lo, hi); // a 64-point transform has at most 32 inputs, 16 of them even
        } else idct16(T, 16, cnt);
    } else {
        if (n == 6) idct64_odd(T, 16, cnt);
        else idct32_odd(T, 16, cnt);
    }
}

template <int n> AV1B_DEV void row_half(const int16_t* coef_row, int tw_nz, bool rect, int half, int* X)
{
    constexpr int M = 1 << (n - 1);
    const int cnt = (tw_nz + 1 - half) >> 1; // x[2k + half], k < cnt, may be non-zero (cnt <= 16)
    int T[M];
    AV1B_UNROLL
    for (int k = 0; k < M; k++) T[k] = (k < 16 && k < cnt) ? (int)coef_row[2 * k + half] : 0;
    if (rect) {
        AV1B_UNROLL
        for (int k = 0; k < 16; k++) T[k] = (T[k] * 2896 + 2048) >> 12;
    }
    dct_half<n>(T, half, cnt);
    AV1B_UNROLL
    for (int k = 0; k < M; k++) X[k] = T[k];
}

template <int n> AV1B_DEV void col_half(const int16_t* tmp_col, int tmp_stride, int nz_rows, int half, int* X)
{
    constexpr int M = 1 << (n - 1);
    const int cnt = (nz_rows + 1 - half) >> 1;
    int T[M];
    AV1B_UNROLL
    for (int k = 0; k < M; k++) T[k] = (k < 16 && k < cnt) ? (int)tmp_col[(2 * k + half) * tmp_stride] : 0;
    dct_half<n>(T, half, cnt);
    AV1B_UNROLL
    for (int k = 0; k < M; k++) X[k] = T[k];
}

// 32-point identity row / column (the only non-DCT transform of that length)
AV1B_DEV void row_identity32(const int16_t* coef_row, int tw_nz, int16_t* tmp_row, bool rect, int row_shift)
{
    AV1B_UNROLL
    for (int j = 0; j < 32; j++) {
        int v = (j < tw_nz) ? (int)coef_row[j] : 0;
        if (rect) v = (v * 2896 + 2048) >> 12;
        tmp_row[j] = (int16_t)clip3(-32768, 32767, round2(v * 4, row_shift));
    }
}
AV1B_DEV void col_identity32(const int16_t* tmp_col, int tmp_stride, int nz_rows, int16_t* out_col, int out_stride, bool fud, int col_shift)
{
    AV1B_UNROLL
    for (int i = 0; i < 32; i++) {
        const int v = (i < nz_rows) ? (int)tmp_col[i * tmp_stride] * 4 : 0;
        out_col[(fud ? (31 - i) : i) * out_stride] = (int16_t)clip3(-32768, 32767, round2(v, col_shift));
    }
}

}  // namespace itx
