// context.cuh -- the device-resident channel (shared by context.cu and slab.cu).  Internal header.
#pragma once
#include "common.cuh"
#include "ops_csr.cuh"
#include "sort.cuh"
#include "tile.cuh"

#include <algorithm>
#include <cmath>

using namespace sphb;

// x-slab of a sharded channel; all zero for an unsharded context
struct SlabState {
    bool on = false;
    int col_lo = 0, col_hi = 0;   // owned global cell columns [col_lo, col_hi)
    int halo = 0;                 // halo width in cell columns
    int ncx_global = 0;
    int n_owned = 0, n_halo = 0;  // fluid rows: owned first, halo copies after
    int n_wall = 0;               // wall rows of the slab and its halo (static)
    int nf_cap = 0;               // fluid rows the buffers can hold
    bool squeezed = false;        // a compaction ran since the last pack: unpack appends after its stayers
    DevBuf w_pos, w_mass, w_wall_vel;   // the static wall store
    DevBuf pid, a_pid;            // int[nf_cap] global identity of each fluid row
    DevBuf sort_flag, sort_ptr;   // int[entry capacity + 1] scratch of the cell-ordered compaction
    DevBuf counts;                // int[16]: block-scan totals / new row counts (device)
    DevBuf block_tot;             // int[5 * blocks]
    int *h_counts = nullptr;      // pinned mirror
    // peer-to-peer exchange (slab.cu, "mailboxes"): this rank's receive area, written by the neighbours' pack kernels
    // over NVLink, and the neighbours' areas this rank's pack kernel writes into
    static constexpr int kMaxWorld = 64;
    double *mail = nullptr;       // own mailbox (cudaMalloc: exportable as a CUDA IPC handle)
    size_t mail_doubles = 0;
    int p2p_rank = 0, p2p_world = 0, p2p_rec_cap = 0;
    double *peer[kMaxWorld] = {}; // mailbox of every rank as seen from this device (own one included)
    bool peer_opened[kMaxWorld] = {};   // mapped with cudaIpcOpenMemHandle (to be closed)
    DevBuf p2p_state;             // long long seq; unsigned ticket; double red[3] (max|v|, tau sums of the whole channel)
};

struct sphb_ctx {
    sphb_params prm{};
    PhysParams P{};
    int device = 0;
    cudaStream_t st = nullptr;
    cudaStream_t own_st = nullptr;   // the context's own stream while `st` is a stream of the caller
    cudaStream_t aux_st = nullptr;   // second stream of sphb_ctx_step_host (copies that overlap the list build)
    cudaEvent_t aux_ev = nullptr;
    // state (+ alternates for the cell-sort permutation)
    DevBuf pos, vel, vel_new, wall_vel, force, force_prior, rho, p, drho_dt, mass, Vol, B;
    DevBuf a_pos, a_vel, a_wall_vel, a_force, a_force_prior, a_rho, a_p, a_drho_dt, a_mass, a_Vol, a_B;
    DevBuf scalars, shear_partials, stage, prof, rp_half;
    DevBuf e_vel, e_w, e_pV, e_b1, e_b2, e_vn, e_rm;   // entry-ordered mirrors (StateView), nw.ell_stride double2 each
    NeighborWork nw;
    SlabState slab;
    SortWork sw;
    bool have_list = false;
    bool list_sized = false;     // the ELL has been grown to the longest row seen at the first build
    bool planned = false;
    bool mirrors_stale = false;  // rows were changed behind a standing list (upload without positions): StateView::e_* lag
    bool pv_valid = false;       // a density pass has written the (p, V) mirror for the current list
    double y_lo = 0, y_hi = 0;
    long long h_step = 0;        // host mirror of the step counter (sort cadence, m:272)
    long long launches = 0;
    // pinned host mirror for status read-back
    DevScalars *h_scalars = nullptr;
    int *h_ints = nullptr;
    // phase timing
    bool timing = false;
    cudaEvent_t ev[8]{};
    sphb_timing last_timing{};
    Prof kprof;                  // per-kernel event timing (sphb_ctx_kernel_profile)
    // cached graphs for the dual-criteria schedule: [0] without, [1] with cell sort
    cudaGraphExec_t graph[2]{nullptr, nullptr};
    int graph_inner = -1;
    long long graph_launches = 0;
    long long outer_count = 0;
    // error flags of graph-replayed runs: copied to pinned memory behind every sphb_ctx_run_dual, looked at (without
    // blocking) at the next call and (blocking) by sphb_ctx_sync / status / download
    cudaEvent_t flags_ev = nullptr;
    bool flags_pending = false;
    unsigned watch_skip = 0;
    int last_inner = 1;          // substeps per outer step of the last run_dual (to redo the steps a poisoned run skipped)
    double last_remain = 1.0e300;   // remaining-time cut of the last single step (m:259), for the same purpose

    StateView view() const
    {
        StateView S;
        S.pos = pos.as<double2>();
        S.vel = vel.as<double2>();
        S.vel_new = vel_new.as<double2>();
        S.wall_vel = wall_vel.as<double2>();
        S.force = force.as<double2>();
        S.force_prior = force_prior.as<double2>();
        S.rho = rho.as<double>();
        S.p = p.as<double>();
        S.drho_dt = drho_dt.as<double>();
        S.mass = mass.as<double>();
        S.Vol = Vol.as<double>();
        S.B = B.as<Mat2>();
        S.rp_half = rp_half.as<double2>();
        S.snap = nw.snap.as<double2>();
        S.grid = nw.grid.as<GridDesc>();
        S.ent_key = nw.ent_key.as<int>();
        S.ent_pos = nw.ent_pos.as<double2>();
        S.ent_of = nw.ent_of.as<int>();
        S.tiles = nw.tiles.as<TileDesc>();
        S.ell_cnt = nw.ell_cnt.as<int>();
        S.ell_q = nw.ell_q.as<uint4>();
        S.e_vel = e_vel.as<double2>();
        S.e_w = e_w.as<double2>();
        S.e_pV = e_pV.as<double2>();
        S.e_b1 = e_b1.as<double2>();
        S.e_b2 = e_b2.as<double2>();
        S.e_vn = e_vn.as<double2>();
        S.e_rm = e_rm.as<double2>();
        S.ghost_ent = nw.ghost_ent.as<int>();
        S.flags = nw.flags.as<int>();
        S.ell_stride = nw.ell_stride;
        S.ell_cap = nw.ell_cap;
        S.tile_cap = nw.tile_cap;
        S.tile_w = nw.tile_w;
        S.ncx = nw.slab ? nw.ncx_local : std::max(1, (int)std::ceil(prm.DL / (2.0 * prm.h)));   // nb:254
        S.n_tiles = nw.n_tiles;
        return S;
    }
    DevScalars *sc() const { return scalars.as<DevScalars>(); }
};


int ctx_bind_device(sphb_ctx *c);
int ctx_plan_lists(sphb_ctx *c, double y_lo, double y_hi);
void ctx_drop_graphs(sphb_ctx *c);
int ctx_do_neighbor(sphb_ctx *c);
int ctx_fresh_mirrors(sphb_ctx *c, bool need_pv);   // makes StateView::e_* current (no-op inside a step)
int ctx_do_shear(sphb_ctx *c);
int ctx_flags_poll(sphb_ctx *c, bool block);          // device error flags behind asynchronous runs (context.cu)
int ctx_flags_watch(sphb_ctx *c);
int ctx_do_sort(sphb_ctx *c);                       // m:272-278 (swaps the state buffers with their alternates)
int ctx_volumes_from_mass(sphb_ctx *c);             // Vol = m / rho for every row (m:120)
int ctx_step_physics(sphb_ctx *c, double remain);   // m:254-270: density + KGC, forces, continuity
