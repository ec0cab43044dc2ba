// imrcd_fit.cuh -- the tree fit shared by the Morton build and the refit (imrcd_fit.cu).
#pragma once
#include <stdint.h>
#include <vector>

#define FIT_T 128u          // triangles per treelet (at most)
#define FIT_R 128u          // records per treelet (at most)
#define FIT_CHAIN 96u       // ancestors above a treelet kept in its chain row (deeper ones are walked by pointer)

// Topology of one tree record as the fit needs it; one per arena record (parallel to d_recs), ARENA indices throughout.
//   first .. last   the record's triangles (leaf order is contiguous per node)        split   last triangle of the left child
//   parent          0xffffffff for a root                                              child   left child (right = child + 1), inner only
//   kind            0 inner, 1 leaf, 2 the padding record beside a root
//   n_sub           records of the subtree, this one included
//   desc            first of the subtree's OTHER records when they are contiguous in the arena (Morton builds: n_sub - 1 records from here),
//                   else 0xffffffff
struct FitRec { uint32_t first, last, split, parent, child, kind, desc, n_sub; };
// a subtree small enough to be fitted by one block in shared memory
#define FIT_SMALL(n_tri, n_sub) ((n_tri) <= FIT_T && (n_sub) <= FIT_R)

struct FitSeg { uint32_t rec_base, n_rec, tri_base, n_tri; double origin[3]; };      // one tree of a fit call; origin is filled on the device

struct FitCounters { uint32_t n_troot, n_upper, n_slots, pad; };

struct imrcd_ctx;
struct MeshDev;
int imr_fit_reserve(imrcd_ctx* ctx, uint64_t n_rec_total);
int imr_fit_plan_from_records(imrcd_ctx* ctx, const MeshDev& md);
int imr_fit_prepare(imrcd_ctx* ctx, const std::vector<FitSeg>& segs, uint64_t topology_key);
int imr_fit_launch(imrcd_ctx* ctx, bool write_links, const uint32_t* bounds, bool classified);
// the lists a builder that classifies its records itself (k_assign) appends to; valid after imr_fit_prepare
struct FitLists { FitCounters* cnt; uint2* troots; uint32_t* uppers; uint32_t* slot_of; };
FitLists imr_fit_lists(imrcd_ctx* ctx);
int imr_fit_reset_counters(imrcd_ctx* ctx);
