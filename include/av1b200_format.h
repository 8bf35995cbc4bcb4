/*
 * av1b200_format.h -- the per-frame command buffer the host front end emits and the
 * sm_100a reconstruction / in-loop-filter engine consumes.
 *
 * One frame = one contiguous byte buffer (written into a pinned host ring slot, copied to
 * HBM with a single cudaMemcpyAsync).  It starts with an Av1bFrameHdr; every other section
 * is addressed by a byte offset from the start of the buffer and is 16-byte aligned.
 *
 * The sections restate, in flat arrays, exactly what the reference's decode() tree walk
 * reads from its parse-time objects:
 *   reference object (file:line)                      -> section
 *   Tile::m_sbs raster (decoder/Tile.cpp:172-181)     -> Av1bSb[]      one per superblock
 *   TransformBlock::decode (TransformBlock.cpp:2376)  -> Av1bOp[]      ordered pixel ops
 *   Block::compute_prediction (Block.cpp:100-174)     -> Av1bInterBlk[] + Av1bIpu[]
 *   Block palette / CfL / inter-intra / warp state    -> Av1bBlkAux[]
 *   TransformBlock::Dequant (TransformBlock.cpp:2266) -> int16 coefficient arena
 *   ModeInfoBlock fields read by LoopFilter/Cdef      -> Av1bLfMi[] + cdef8[]
 *   LoopRestorationpParams per-unit state             -> Av1bLrUnit[]
 *
 * Plain C, no CUDA or torch types: this header is part of the C ABI.
 */
#ifndef AV1B200_FORMAT_H_
#define AV1B200_FORMAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AV1B_MAGIC 0x42315641u /* "AV1B" */
#define AV1B_FORMAT_VERSION 4

/* ---- Av1bOp.kind ------------------------------------------------------------------- */
enum {
    AV1B_OP_INTER_RES = 0,  /* TB of an inter-intra / intrabc block: frame += residual (other inter
                               blocks get their residual added by the inter pass itself)        */
    AV1B_OP_INTRA = 1,      /* intra TB: predict from neighbours, + residual                      */
    AV1B_OP_PALETTE = 2,    /* palette TB: paint colour map, + residual                           */
    AV1B_OP_INTERINTRA = 3, /* whole plane-block: intra predict and blend over the inter pred    */
    AV1B_OP_INTRABC = 4     /* whole plane-block: motion-compensate from the CURRENT frame        */
};

/* ---- Av1bOp.flags ------------------------------------------------------------------ */
#define AV1B_OPF_HAVE_LEFT 0x01
#define AV1B_OPF_HAVE_ABOVE 0x02
#define AV1B_OPF_HAVE_ABOVE_RIGHT 0x04
#define AV1B_OPF_HAVE_BELOW_LEFT 0x08
#define AV1B_OPF_EDGE_SMOOTH 0x10  /* get_filter_type(): a neighbour uses a SMOOTH mode */
#define AV1B_OPF_HAS_RESID 0x20    /* eob > 0: residual present at res_off                */
#define AV1B_OPF_CFL 0x40          /* chroma-from-luma on top of the DC prediction         */
#define AV1B_OPF_FILTER_INTRA 0x80 /* recursive filter-intra (luma only)                   */

/* One ordered pixel operation of the dependent ("wavefront") reconstruction pass. 32 bytes. */
typedef struct Av1bOp {
    uint16_t x, y;      /* top-left in plane samples                                        */
    uint8_t plane;      /* 0,1,2                                                            */
    uint8_t kind;       /* AV1B_OP_*                                                        */
    uint8_t tx_size;    /* TX_SIZE enum; INTERINTRA/INTRABC: log2w | (log2h << 4)           */
    uint8_t tx_type;    /* TX_TYPE enum (PlaneTxType)                                       */
    uint8_t mode;       /* intra prediction mode actually run (CfL -> DC_PRED)              */
    int8_t angle_delta; /* AngleDeltaY / AngleDeltaUV                                       */
    uint8_t flags;      /* AV1B_OPF_*                                                       */
    uint8_t fi_mode;    /* bits 0-2: filter_intra_mode.  Intra ops without filter-intra: bits 3-4
                           log2 of the number of row strips the block is split into, bits 5-7 the
                           strip this op predicts (one op per strip, consecutive; 0 = whole block) */
    int8_t cfl_alpha;   /* CflAlphaU / CflAlphaV                                            */
    uint8_t nz_rows;    /* number of leading coefficient rows that may be non-zero (<=32)  */
    uint8_t nz_cols;    /* number of leading coefficient columns that may be non-zero      */
    uint8_t lossless;   /* 1: Walsh-Hadamard path (Block::Lossless)                         */
    uint32_t coef_off;  /* int16 index into the coefficient arena (tw*th values, row-major) */
    uint32_t res_off;   /* stage-level ITX test mode: int16 index into the compact residual arena.
                           Full submits (residuals go to frame-layout int16 planes): dependency
                           level of the op inside its superblock in the low 16 bits, number of ops
                           left in that level (this one included) in the high 16 bits            */
    uint32_t aux;       /* Av1bBlkAux index (palette, inter-intra) or Av1bIpu index (intrabc) */
    uint16_t max_luma_w, max_luma_h; /* CfL: Block::MaxLumaW/H at this TB                   */
} Av1bOp;

/* Superblock entry: its ops are [first_op, first_op + n_ops), sorted by dependency level (levels
 * count from 1).  Indexed by frame-raster SB.
 *
 * The remaining fields let neighbouring superblocks OVERLAP in the wavefront: a superblock does not
 * wait for all of its left / above-right neighbour before its first op, only -- level by level --
 * for the part of the neighbour's border it is about to read, and it announces the halves of its
 * own right column / bottom row as soon as no later op writes them.  ("half" = sb/2 samples.)
 *   wait_l1  first level that reads the left superblock's right column above its middle
 *   wait_l2  first level that reads it below the middle
 *   wait_a1  first level that reads the above-right superblock's bottom row left of its middle
 *   wait_a2  first level that reads it right of the middle
 *            (0 = before the first level, 0xFF = never)
 *   pub_r1   level after which the upper half of this superblock's right column is final
 *   pub_b1   level after which the left half of its bottom row is final   (0 = only at the end)
 * The superblocks above and above-left are always complete before a superblock starts.  All zero
 * is the conservative setting: the classic two-superblock-lag wavefront. */
typedef struct Av1bSb {
    uint32_t first_op;
    uint32_t n_ops;
    uint8_t wait_l1, wait_l2, wait_a1, wait_a2;
    uint8_t pub_r1, pub_b1;
    uint8_t pad[2];
} Av1bSb;

/* ---- inter prediction ---------------------------------------------------------------- */
enum { AV1B_IPU_PRED = 0, AV1B_IPU_OBMC_ABOVE = 1, AV1B_IPU_OBMC_LEFT = 2 };
/* comp_type values follow the reference COMPOUND_TYPE enum (aom/enums.h:474-481) */
enum {
    AV1B_COMP_WEDGE = 0,
    AV1B_COMP_DIFFWTD = 1,
    AV1B_COMP_AVERAGE = 2,
    AV1B_COMP_INTRA = 3,
    AV1B_COMP_DISTANCE = 4
};
#define AV1B_IPUF_COMPOUND 0x01   /* two reference lists                                    */
#define AV1B_IPUF_INTERINTRA 0x02 /* block is inter-intra: write Clip1(pred) only           */
#define AV1B_IPUF_INTRABC 0x04    /* reference is the current frame (run inside wavefront)  */
#define AV1B_IPUF_FAST 0x08       /* unit of an AV1B_IBF_FAST block                          */
#define AV1B_IPUF_ADD_RES 0x10    /* fast / independent unit: add the residual planes while writing */
#define AV1B_IPUF_INDEP 0x20      /* unit of an AV1B_IBF_FAST block that needs the general predictor */

/* One predict_inter() call (reference decoder/InterPredict.cpp:962) or one OBMC strip. 32 B */
typedef struct Av1bIpu {
    uint16_t x, y;      /* destination top-left in plane samples */
    uint8_t w, h;       /* size in plane samples (<=128)         */
    uint8_t plane;
    uint8_t kind;       /* AV1B_IPU_*                            */
    int16_t mv[2][2];   /* [refList][0=row,1=col], 1/8 luma pel  */
    int8_t ref_slot[2]; /* frame-store slot per list             */
    uint8_t ref_frame[2]; /* RefFrame (1..7) per list: ref dims and global-motion params    */
    uint8_t filt[2];    /* InterpFilters[0] (vertical), [1] (horizontal) of the candidate MI */
    uint8_t warp[2];    /* per list: 0 none, 1 local warp, 2 global warp (size test on device) */
    uint8_t flags;      /* AV1B_IPUF_*                           */
    uint8_t comp_type;  /* AV1B_COMP_*                           */
    uint8_t fwd_w, bck_w; /* distance weights                    */
    uint32_t aux;       /* Av1bBlkAux index (warp params, wedge, mask)                      */
} Av1bIpu;

#define AV1B_IBF_HAS_CHROMA 0x01   /* the block carries chroma (Block::HasChroma)                   */
#define AV1B_IBF_ADD_RESIDUAL 0x02 /* add the residual planes over the block after prediction        */
/* The units of the block are independent of each other (all of kind PRED: no OBMC strips; no
 * diff-weighted compound mask shared between the planes): they are handled one by one, by the fast
 * translational kernel (AV1B_IPUF_FAST: no warp, no mask, w >= 4, compound average / distance at
 * most) or by the general predictor with a warp per unit (AV1B_IPUF_INDEP), either of which also
 * adds the residual (AV1B_IPUF_ADD_RES).  Optional: a producer that never sets it gets the
 * block-at-a-time kernel for every block. */
#define AV1B_IBF_FAST 0x04

/* One inter block = one CTA work item of the independent inter pass. 24 B */
typedef struct Av1bInterBlk {
    uint32_t first_ipu;
    uint16_t n_ipu;
    uint16_t flags;      /* AV1B_IBF_*                                   */
    uint16_t x, y;       /* luma rectangle (samples)                     */
    uint16_t cx, cy;     /* chroma rectangle origin (chroma samples)     */
    uint8_t bw, bh;      /* luma size; 0 means 256 is never needed (<=128) */
    uint8_t cw, ch;      /* chroma size                                  */
    uint32_t pad;
} Av1bInterBlk;

/* Rarely-needed per-block side data. 96 B */
typedef struct Av1bBlkAux {
    int32_t warp_params[6];  /* LocalWarpParams                                              */
    int16_t warp_abgd[4];    /* alpha, beta, gamma, delta from setupShear()                  */
    uint8_t mi_size;         /* BLOCK_SIZE                                                   */
    uint8_t interintra_mode; /* II_*                                                         */
    uint8_t wedge_interintra;
    uint8_t wedge_index;
    uint8_t wedge_sign;
    uint8_t mask_type;
    uint8_t pal_size_y, pal_size_uv;
    uint8_t pal_colors[3][8];
    uint32_t pal_map_off[2]; /* byte offsets into the palette arena (Y map, UV map)          */
    uint16_t pal_map_stride[2];
    uint16_t base_x[2], base_y[2]; /* block origin in plane samples (luma, chroma)           */
    uint8_t pad[8];
} Av1bBlkAux;

/* ---- in-loop filter side data -------------------------------------------------------- */
/* Per 4x4 mode-info record read by the deblocking filter (LoopFilter.cpp:85-126,301-359). 8 B */
typedef struct Av1bLfMi {
    uint8_t mi_size; /* BLOCK_SIZE */
    uint8_t flags;   /* b0 Skip, b1 modeType (YMode>=NEARESTMV && not GLOBAL*), b2..4 RefFrames[0] clipped to >=0 */
    uint16_t tx;     /* LoopfilterTxSizes: Y | U<<5 | V<<10 */
    int8_t delta_lf[4];
} Av1bLfMi;

/* Loop-restoration unit (LoopRestorationpParams LrType/LrWiener/LrSgrSet/LrSgrXqd). 12 B */
typedef struct Av1bLrUnit {
    uint8_t type;       /* RESTORE_NONE=0, RESTORE_WIENER=1, RESTORE_SGRPROJ=2 */
    uint8_t sgr_set;
    int8_t sgr_xqd[2];
    int8_t wiener[2][3]; /* [0]=vertical pass coeffs, [1]=horizontal */
    uint8_t pad[2];
} Av1bLrUnit;

typedef struct Av1bLoopFilterParams {
    uint8_t level[4];
    uint8_t sharpness;
    uint8_t delta_enabled;
    uint8_t delta_lf_multi;
    uint8_t pad;
    int8_t ref_deltas[8];
    int8_t mode_deltas[2];
    uint8_t pad2[6];
} Av1bLoopFilterParams;

typedef struct Av1bCdefParams {
    uint8_t enabled; /* 0: every cdef8 entry is 0xFF, stage skipped */
    uint8_t damping;
    uint8_t pad[2];
    uint8_t y_pri[8], y_sec[8], uv_pri[8], uv_sec[8];
} Av1bCdefParams;

typedef struct Av1bLrParams {
    uint8_t uses_lr;
    uint8_t frame_type[3]; /* FrameRestorationType per plane (0 = none) */
    uint16_t unit_size[3];
    uint16_t unit_rows[3];
    uint16_t unit_cols[3];
    uint16_t pad;
    uint32_t unit_first[3]; /* index of the plane's first Av1bLrUnit (row-major after it) */
} Av1bLrParams;

/* Frame header: first bytes of the command buffer. */
typedef struct Av1bFrameHdr {
    uint32_t magic;
    uint32_t version;
    uint32_t total_bytes;
    uint16_t frame_w, frame_h; /* FrameWidth / FrameHeight (== UpscaledWidth: no super-res) */
    uint16_t mi_cols, mi_rows;
    uint16_t sb_cols, sb_rows;
    uint8_t sb_log2;           /* 6 or 7 */
    uint8_t enable_intra_edge_filter;
    uint8_t frame_is_intra;
    uint8_t allow_intrabc;
    /* references, indexed by RefFrame 1..7 ([0] = current frame for intrabc) */
    int8_t ref_slot[8];
    uint16_t ref_w[8], ref_h[8]; /* RefUpscaledWidth / RefFrameHeight */
    int32_t gm_params[8][6];
    int16_t gm_abgd[8][4];
    uint8_t gm_warp_ok[8];       /* GmType > TRANSLATION && setupShear() valid */
    /* sections */
    uint32_t off_sb, n_sb;
    uint32_t off_ops, n_ops;
    uint32_t off_itx, n_itx;   /* uint32 op indices that carry a residual (inverse-transform work list),
                                  sorted by size class: max(w,h) = 4 | 8 | 16 | >= 32                  */
    uint32_t itx_class_end[4]; /* end index (exclusive) of each size class inside the list              */
    uint32_t off_iblk, n_iblk;
    uint32_t off_ipu, n_ipu;
    uint32_t off_aux, n_aux;
    uint32_t off_coef, n_coef; /* int16 count */
    uint32_t n_res;            /* int16 count of the device-side residual arena */
    uint32_t off_pal, n_pal;   /* bytes */
    uint32_t off_lfmi;         /* Av1bLfMi[mi_rows * mi_cols] */
    uint32_t off_cdef8;        /* uint8[(mi_rows/2) * (mi_cols/2)]: preset index, or 0xFF = leave 8x8 untouched */
    uint32_t off_lru, n_lru;   /* Av1bLrUnit[] */
    Av1bLoopFilterParams lf;
    Av1bCdefParams cdef;
    Av1bLrParams lr;
} Av1bFrameHdr;

#ifdef __cplusplus
}
#endif
#endif /* AV1B200_FORMAT_H_ */
