// The device-resident simulation object behind atlj_sim (internal).
#pragma once
#include <vector>

#include "common.cuh"
#include "pair_math.cuh"

// slab decomposition state (mg.cu)
// Per-atom arrays of a slab rank: [0, own_cap) owned atoms (cell order), [own_cap, own_cap + halo_cap) the left
// neighbour's last layer (local layer 0), [own_cap + halo_cap, cap) the right neighbour's first layer (local layer
// nxl - 1).  The halo regions of r[0], r[1] and the halo rows of cell_start are WRITTEN BY THE NEIGHBOURS (peer memory,
// NVLink); only migrating atoms (r, v, id: a few hundred per face and step) go through the message arena.
struct MgState {
  bool on = false;
  int rank = 0, world = 1;
  int64_t mig_cap = 0, halo_cap = 0, own_cap = 0;  // atoms
  size_t mig_bytes = 0, arena_bytes = 0;
  unsigned char *arena = nullptr;  // my receive buffers: migration [from left|right][parity], then 2 halo headers
  // the neighbours' memory, mapped into this process (CUDA IPC) or plain pointers (ranks emulated in one process)
  unsigned char *peer_arena[2] = {nullptr, nullptr};
  double *peer_r[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [left|right][buffer]
  int *peer_cs[2] = {nullptr, nullptr};
  double *peer_f[2] = {nullptr, nullptr};  // the neighbours' force arrays (hessian on slabs: forces of the halo atoms)
  int peer_nxl[2] = {0, 0};
  void *ipc_mapped[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // to be closed
  int n_ipc_mapped = 0;
  bool sorted = false;      // the arrays are in cell order and cell_start describes them (false after an upload)
  atlj::RecTail tail;       // phase 1 -> phase 2: the record the scan launch has to finish (nrows = 0: none)
  long long step = 0;                                   // exchange counter: parity selects the buffer, step+1 = stamp
  int *cnt_dev = nullptr;   // [0] atoms held after the arrivals were appended, [1] owned atoms (after the sort),
                            // [4..6] migration slot counters + CTAs done, [8], [9] halo pack CTAs done
  int *cnt_host = nullptr;  // pinned mirror of cnt_dev[1]
  void *comm = nullptr;     // ncclComm_t (all-reduce of the step records)
  unsigned long long *trace = nullptr;  // ATLJ_MG_TRACE=1: %globaltimer marks between the launches of a step
  unsigned long long timeout_ns = 30000000000ull;  // a neighbour's message: give up after this long (ATLJ_PEER_TIMEOUT_MS)
};

struct atlj_sim {
  cudaStream_t st = nullptr;
  int64_t cap = 0;  // atom capacity of every per-atom buffer
  bool ipc_safe = false;     // r[0], r[1], f, cell_start are allocations of their own (exportable with CUDA IPC)
  size_t r_alloc_bytes = 0, cs_alloc_bytes = 0;  // their sizes when ipc_safe (a signature sits in the last 256 bytes)
  size_t f_alloc_bytes = 0;                      // ... and of the force array
  int64_t n = 0;    // owned atoms
  int64_t n_total = 0;       // atoms of the whole system (all ranks); indexes injected noise by global id
  long long step_count = 0;  // steps taken so far (counter of the device RNG)
  bool configured = false;
  bool defer_fin = false;   // set by a loop before sim_force: leave the totals' final reduction to my next kernel ...
  atlj::FinTail fin;        // ... which finds it here (nb = 0: sim_force has launched finalize_kernel itself)
  bool forces_valid = false;  // s->f belongs to the current positions (cleared by upload / configure, set by sim_force)
  bool use_cells = false;  // sc >= 4: link-cell route; otherwise all pairs
  atlj::Geom g{};
  atlj::Phys p{};
  // state, double-buffered because the cell sort re-orders atoms every force evaluation
  double *r[2] = {nullptr, nullptr};
  double *v[2] = {nullptr, nullptr};
  int *id[2] = {nullptr, nullptr};
  int cur = 0;
  double *f = nullptr;  // forces in the order of r[cur]
  atlj::CellBufs cb{};
  int ncell_alloc = 0;
  double *partials = nullptr;  // [partials_rows][8]
  int partials_rows = 0;
  int pair_kernel = 0;  // 0 = tile2 (LJ force) / tile1 (WCA, hessian); 1 = tile1 everywhere
  int *tile_tab = nullptr;     // work-item table of pair_tile2
  int tile_tab_ints = 0;
  uint2 *masks = nullptr;      // pair_tile2 mask words (allocated with the first hessian request)
  double *vf_partials = nullptr;  // [<= 4096][4] partial {sum v^2, sum f^2} of the integration kernels
  bool unfused = false;
  bool no_graph = false;
  // replayed NVE loop: two captured steps (api.cu nve_graph_prepare)
  cudaGraphExec_t nve_graph = nullptr;
  int nve_graph_launches = 0;       // kernel launches one replay stands for
  unsigned char nve_graph_key[160] = {0};
  cudaStream_t st_cap = nullptr, st_side = nullptr;  // capture stream and its fork (item table beside the sort)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double *scal = nullptr;      // 64 device scalars (thermostat state, reduction results)
  double *rec_dev = nullptr;   // per-step records
  int64_t rec_cap = 0;
  double *noise_dev = nullptr;  // injected noise (tests)
  int64_t noise_cap = 0;
  MgState mg;
  // host-buffer step: copy engine stream + events so that r, id go home while the pair kernel runs
  cudaStream_t st_copy = nullptr;
  cudaEvent_t ev_sorted = nullptr, ev_copied = nullptr;
  // timing
  cudaEvent_t sw0 = nullptr, sw1 = nullptr;  // stopwatch (atlj_sim_timer_*)
  bool timing = false;
  struct Ev {
    cudaEvent_t e0, e1;
    int cat;
  };
  std::vector<Ev> ev_open2;  // brackets open on the second stream
  std::vector<cudaEvent_t> ev_free;
  std::vector<Ev> ev_open, ev_done;
};

namespace atlj {
struct Timer {
  // cat: 0 pair, 1 cells, 2 integrate, 3 other, 4 migration exchange, 5 halo pack, 6 halo exchange + unpack,
  // 7 boundary pair; st = the stream the bracketed work runs on (default: the sim's stream)
  static void begin(atlj_sim *s, int cat, cudaStream_t st = nullptr, bool other_stream = false);
  static void end(atlj_sim *s, cudaStream_t st = nullptr, bool other_stream = false);
};
// after_sort (may be null): recorded on the sim's stream when the sorted r, v, id of this evaluation are final
int sim_force(atlj_sim *s, double strain, bool with_hessian, double *rec_dev, bool check_index_allpairs,
              cudaEvent_t after_sort = nullptr);
int sim_ensure_forces(atlj_sim *s, double strain);  // evaluates s->f when it is not current (after an upload)
int sim_check_flags(atlj_sim *s);
double sim_per_cell(const atlj_sim *s);
bool sim_tile2(const atlj_sim *s);
int sim_ensure_masks(atlj_sim *s);  // mask words of pair_tile2 (kept by the force kernel for the hessian kernel)
int sim_copy_rec(atlj_sim *s, double *rec_host, int nsteps);
int sim_ensure_rec(atlj_sim *s, int64_t nrec);
void sim_mg_destroy(atlj_sim *s);
// slab building blocks for the integrators of dynamics.cu (mg.cu)
const int *mg_n_dev(const atlj_sim *s);  // owned atoms, on the device
long long mg_own_cap(const atlj_sim *s);
int mg_begin(atlj_sim *s);               // host -> device counters at the start of a run
int mg_end(atlj_sim *s);                 // device -> host owned count at its end
int mg_assign_exchange_force(atlj_sim *s, double *rec, bool first);  // bin, migrate, sort, halo, pair kernel, totals
int mg_phase1_assign(atlj_sim *s, bool first);  // emulation: cell assignment + migration of a step moved elsewhere
int mg_allreduce(atlj_sim *s, double *dev, size_t count);
bool mg_connected(const atlj_sim *s);
}  // namespace atlj
