// imrcd_fit.cuh -- the tree fit shared by the Morton build and the refit (imrcd_fit.cu).
#pragma once
#include <stdint.h>
#include <vector>

#define FIT_T 128u          // triangles per treelet (at most)
#define FIT_R 256u          // records per treelet (at most 2 * FIT_T - 1: leaves hold >= 1 triangle)

// Topology of one tree record as the fit needs it; one per arena record (parallel to d_recs), ARENA indices throughout.
//   first .. last   the record's triangles (leaf order is contiguous per node)        split   last triangle of the left child
//   parent          0xffffffff for a root                                              child   left child (right = child + 1), inner only
//   kind            0 inner, 1 leaf, 2 the padding record beside a root
struct FitRec { uint32_t first, last, split, parent, child, kind, pad0, pad1; };

struct FitSeg { uint32_t rec_base, n_rec, tri_base, n_tri; double origin[3]; };      // one tree of a fit call; origin is filled on the device

struct FitCounters { uint32_t n_troot, n_upper, n_slots, pad; };

struct imrcd_ctx;
struct MeshDev;
int imr_fit_reserve(imrcd_ctx* ctx, uint64_t n_rec_total);
int imr_fit_plan_from_records(imrcd_ctx* ctx, const MeshDev& md);
int imr_fit_prepare(imrcd_ctx* ctx, const std::vector<FitSeg>& segs);
int imr_fit_launch(imrcd_ctx* ctx, bool write_links, const uint32_t* bounds, const uint32_t* n_inner_dev);
