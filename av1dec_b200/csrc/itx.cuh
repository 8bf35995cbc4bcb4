// itx.cuh -- AV1 inverse transforms (DCT4..64, ADST4/8/16, identity4..32, WHT4), bit-exact.
//
// Follows the AV1 specification section 7.13 (inverse transform process) as the reference
// transcribes it in decoder/TransformBlock.cpp:1751-2253 -- same butterfly order, same
// Round2 placement, same intermediate clamps -- because bit-exactness requires the same
// arithmetic.  The structure is different: every 1-D transform is a template on log2(N)
// whose loops unroll completely so the N-point vector lives in registers; the angle tables
// and bit-reversals fold to immediates.
#pragma once
#include "dev.h"
#include "av1_tables.h"

namespace itx {

AV1B_DEV constexpr int brev(int bits, int x)
{
    int t = 0;
    for (int i = 0; i < bits; i++)
        t |= ((x >> i) & 1) << (bits - 1 - i);
    return t;
}

// cos(angle*pi/128) in Q12 (reference cos128(), TransformBlock.cpp:1782)
AV1B_DEV int cos128(int angle)
{
    int a = angle & 255;
    if (a <= 64) return k_cos128[a];
    if (a <= 128) return -k_cos128[128 - a];
    if (a <= 192) return -k_cos128[a - 128];
    return k_cos128[256 - a];
}
AV1B_DEV int sin128(int angle) { return cos128(angle - 64); }

// Butterfly rotation (reference B(), TransformBlock.cpp:1799)
AV1B_DEV void rot(int* T, int a, int b, int angle, bool flip)
{
    int c = cos128(angle), s = sin128(angle);
    int x = T[a] * c - T[b] * s;
    int y = T[a] * s + T[b] * c;
    x = (x + 2048) >> 12;
    y = (y + 2048) >> 12;
    if (!flip) {
        T[a] = x;
        T[b] = y;
    } else {
        T[b] = x;
        T[a] = y;
    }
}

// Hadamard add/sub with clamp to r bits (reference H(), TransformBlock.cpp:1815)
AV1B_DEV void had(int* T, int a, int b, bool flip, int r)
{
    if (flip) {
        int t = a;
        a = b;
        b = t;
    }
    int lo = -(1 << (r - 1)), hi = (1 << (r - 1)) - 1;
    int x = T[a], y = T[b];
    T[a] = clip3(lo, hi, x + y);
    T[b] = clip3(lo, hi, x - y);
}

// Inverse DCT, N = 1 << n (spec 7.13.2.3; reference iDct(), TransformBlock.cpp:1827-1989)
template <int n> AV1B_DEV void idct(int* T, int r)
{
    constexpr int N = 1 << n;
    {
        int c[N];
        AV1B_UNROLL
        for (int i = 0; i < N; i++) c[i] = T[i];
        AV1B_UNROLL
        for (int i = 0; i < N; i++) T[i] = c[brev(n, i)];
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 16; i++) rot(T, 32 + i, 63 - i, 63 - 4 * brev(4, i), false);
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) rot(T, 16 + i, 31 - i, 6 + (brev(3, 7 - i) << 3), false);
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 16; i++) had(T, 32 + i * 2, 33 + i * 2, i & 1, r);
    }
    if (n >= 4) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) rot(T, 8 + i, 15 - i, 12 + (brev(2, 3 - i) << 4), false);
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) had(T, 16 + 2 * i, 17 + 2 * i, i & 1, r);
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 2; j++) rot(T, 62 - i * 4 - j, 33 + i * 4 + j, 60 - 16 * brev(2, i) + 64 * j, true);
        }
    }
    if (n >= 3) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) rot(T, 4 + i, 7 - i, 56 - 32 * i, false);
    }
    if (n >= 4) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) had(T, 8 + 2 * i, 9 + 2 * i, i & 1, r);
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 2; j++) rot(T, 30 - 4 * i - j, 17 + 4 * i + j, 24 + (j << 6) + ((1 - i) << 5), true);
        }
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 2; j++) had(T, 32 + i * 4 + j, 35 + i * 4 - j, i & 1, r);
        }
    }
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) rot(T, 2 * i, 2 * i + 1, 32 + 16 * i, 1 - i);
    if (n >= 3) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) had(T, 4 + 2 * i, 5 + 2 * i, i, r);
    }
    if (n >= 4) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) rot(T, 14 - i, 9 + i, 48 + 64 * i, true);
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 2; j++) had(T, 16 + 4 * i + j, 19 + 4 * i - j, i & 1, r);
        }
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 4; j++) rot(T, 61 - i * 8 - j, 34 + i * 8 + j, 56 - i * 32 + (j >> 1) * 64, true);
        }
    }
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) had(T, i, 3 - i, false, r);
    if (n >= 3) rot(T, 6, 5, 32, true);
    if (n >= 4) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 2; j++) had(T, 8 + 4 * i + j, 11 + 4 * i - j, i, r);
        }
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) rot(T, 29 - i, 18 + i, 48 + (i >> 1) * 64, true);
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 4; j++) had(T, 32 + 8 * i + j, 39 + 8 * i - j, i & 1, r);
        }
    }
    if (n >= 3) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) had(T, i, 7 - i, false, r);
    }
    if (n >= 4) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) rot(T, 13 - i, 10 + i, 32, true);
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 2; i++) {
            AV1B_UNROLL
            for (int j = 0; j < 4; j++) had(T, 16 + i * 8 + j, 23 + i * 8 - j, i, r);
        }
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) rot(T, 59 - i, 36 + i, i < 4 ? 48 : 112, true);
    }
    if (n >= 4) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) had(T, i, 15 - i, false, r);
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 4; i++) rot(T, 27 - i, 20 + i, 32, true);
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) {
            had(T, 32 + i, 47 - i, false, r);
            had(T, 48 + i, 63 - i, true, r);
        }
    }
    if (n >= 5) {
        AV1B_UNROLL
        for (int i = 0; i < 16; i++) had(T, i, 31 - i, false, r);
    }
    if (n == 6) {
        AV1B_UNROLL
        for (int i = 0; i < 8; i++) rot(T, 55 - i, 40 + i, 32, true);
        AV1B_UNROLL
        for (int i = 0; i < 32; i++) had(T, i, 63 - i, false, r);
    }
}

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

template <int n> AV1B_DEV void iadst_in_perm(int* T)
{
    constexpr int N = 1 << n;
    int c[N];
    AV1B_UNROLL
    for (int i = 0; i < N; i++) c[i] = T[i];
    AV1B_UNROLL
    for (int i = 0; i < N; i++) T[i] = c[(i & 1) ? (i - 1) : (N - i - 1)];
}

template <int n> AV1B_DEV void iadst_out_perm(int* T)
{
    constexpr int N = 1 << n;
    int c[N];
    AV1B_UNROLL
    for (int i = 0; i < N; i++) c[i] = T[i];
    AV1B_UNROLL
    for (int i = 0; i < N; i++) {
        int a = (i >> 3) & 1;
        int b = ((i >> 2) & 1) ^ ((i >> 3) & 1);
        int cc = ((i >> 1) & 1) ^ ((i >> 2) & 1);
        int d = (i & 1) ^ ((i >> 1) & 1);
        int idx = ((d << 3) | (cc << 2) | (b << 1) | a) >> (4 - n);
        T[i] = (i & 1) ? -c[idx] : c[idx];
    }
}

// ADST8 (spec 7.13.2.7; reference iAdst8(), TransformBlock.cpp:2053)
AV1B_DEV void iadst8(int* T, int r)
{
    iadst_in_perm<3>(T);
    AV1B_UNROLL
    for (int i = 0; i < 4; i++) rot(T, 2 * i, 2 * i + 1, 60 - 16 * i, true);
    AV1B_UNROLL
    for (int i = 0; i < 4; i++) had(T, i, 4 + i, false, r);
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) rot(T, 4 + 3 * i, 5 + i, 48 - 32 * i, true);
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) {
        AV1B_UNROLL
        for (int j = 0; j < 2; j++) had(T, 4 * j + i, 2 + 4 * j + i, false, r);
    }
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) rot(T, 2 + 4 * i, 3 + 4 * i, 32, true);
    iadst_out_perm<3>(T);
}

// ADST16 (spec 7.13.2.8; reference iAdst16(), TransformBlock.cpp:2075)
AV1B_DEV void iadst16(int* T, int r)
{
    iadst_in_perm<4>(T);
    AV1B_UNROLL
    for (int i = 0; i < 8; i++) rot(T, 2 * i, 2 * i + 1, 62 - 8 * i, true);
    AV1B_UNROLL
    for (int i = 0; i < 8; i++) had(T, i, 8 + i, false, r);
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) {
        rot(T, 8 + 2 * i, 9 + 2 * i, 56 - 32 * i, true);
        rot(T, 13 + 2 * i, 12 + 2 * i, 8 + 32 * i, true);
    }
    AV1B_UNROLL
    for (int i = 0; i < 4; i++) {
        AV1B_UNROLL
        for (int j = 0; j < 2; j++) had(T, 8 * j + i, 4 + 8 * j + i, false, r);
    }
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) {
        AV1B_UNROLL
        for (int j = 0; j < 2; j++) rot(T, 4 + 8 * j + 3 * i, 5 + 8 * j + i, 48 - 32 * i, true);
    }
    AV1B_UNROLL
    for (int i = 0; i < 2; i++) {
        AV1B_UNROLL
        for (int j = 0; j < 4; j++) had(T, 4 * j + i, 2 + 4 * j + i, false, r);
    }
    AV1B_UNROLL
    for (int i = 0; i < 4; i++) rot(T, 2 + 4 * i, 3 + 4 * i, 32, true);
    iadst_out_perm<4>(T);
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

template <int n> AV1B_DEV void run1d(int* T, int kind, int r, int wht_shift)
{
    if (kind == K_DCT) {
        idct<n>(T, r);
    } else if (kind == K_ADST) {
        if (n == 2) iadst4(T);
        else if (n == 3) iadst8(T, r);
        else if (n == 4) iadst16(T, r);
    } else if (kind == K_IDT) {
        if (n <= 5) iidentity<(n <= 5 ? n : 5)>(T);
    } else {
        if (n == 2) iwht4(T, wht_shift);
    }
}

// Row pass of one coefficient row: coef (int16, tw valid entries) -> tmp row (int16, N entries).
template <int n>
AV1B_DEV void row_pass(const int16_t* coef_row, int tw, int16_t* tmp_row, int kind, bool rect, int row_shift)
{
    constexpr int N = 1 << n;
    int T[N];
    AV1B_UNROLL
    for (int j = 0; j < N; j++) T[j] = (j < tw) ? (int)coef_row[j] : 0;
    if (rect) {
        AV1B_UNROLL
        for (int j = 0; j < N; j++) T[j] = (T[j] * 2896 + 2048) >> 12;
    }
    run1d<n>(T, kind, 16, 2);
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
    run1d<n>(T, kind, 16, 0);
    AV1B_UNROLL
    for (int i = 0; i < N; i++) {
        int v = clip3(-32768, 32767, round2(T[i], col_shift));
        out_col[(fud ? (N - 1 - i) : i) * out_stride] = (int16_t)v;
    }
}

}  // namespace itx
