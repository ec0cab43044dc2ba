// imrcd_frame.cuh -- what the translation units of a frame share: imrcd_frame.cu (broad phase, pair setup, orchestration),
// imrcd_traverse.cu (mid phase), imrcd_narrow.cu (narrow phase), imrcd_contacts.cu (contact reduction), imrcd_rays.cu (response).
#pragma once
#include "imrcd_internal.cuh"

#ifndef FULL_MASK
#define FULL_MASK 0xffffffffu
#endif

// contact reduction: size classes with their tables in shared memory - <= 256 hits (128 threads), <= 512 (512 threads), <= 1024 (1024 threads) -
// and the large pairs (more hits), which go through the grid-wide passes k_large_* with their tables in a global scratch
#define PC_S_MAX 256u
#define PC_M1_MAX 512u
#define PC_M_MAX 1024u
#define PC_CLASSES 4
struct PcSlot { double w, cx, cy, cz; };
struct LargeSide { uint32_t n_avg, n_vert, ray_base, cursor; };

// narrow phase (imrcd_narrow.cu)
int imr_narrow_prepare(imrcd_ctx* ctx);
int imr_narrow_launch(imrcd_ctx* ctx, FrameCtl* ctl);
// contact reduction (imrcd_contacts.cu): hit lists, grouping, the per-pair and the grid-wide reductions; everything between the narrow phase and k_finalize
int imr_contacts_prepare(imrcd_ctx* ctx);
int imr_contacts_enqueue(imrcd_ctx* ctx, FrameCtl* ctl);
