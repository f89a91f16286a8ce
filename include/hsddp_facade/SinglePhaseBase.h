// SinglePhaseBase.h — the phase interface MultiPhaseDDP<T> is built on (HSDDPSolver/header/SinglePhaseBase.h:10-87).
// In the reference a phase object owns callbacks and does the per-phase numerics on the host. Here the numerics of all phases run
// inside the CUDA solver, so the interface keeps what the callers of the solver use (dimensions, shooting-state configuration, print)
// and adds the binding a phase has to the plain-data phase deck (which deck, which phase of it) plus the two copies between its
// Trajectory and the solver's packed solution record.
#pragma once
#include <memory>
#include "HSDDP_CPPTypes.h"
#include "HSDDP_CompoundTypes.h"

template <typename> class MultiPhaseDDP;

namespace cafe_facade {
// the deck a problem builder made, shared by the phases cut from it (kept alive as long as one phase refers to it)
struct DeckOwner {
  CafeDeckHandle* h = nullptr;
  int k0 = 0;                  // start offset (in reference rows) the deck was built at; advanced by <Problem>::update()
  long lineage = 0;            // decks that <Problem>::update() derives from one another share it: the solver may then carry state that lives on
                               // the device over the update (relaxed-barrier update counts) instead of starting cold
  static long next_lineage() { static long n = 0; return ++n; }
  ~DeckOwner() { if (h) cafe_deck_free(h); }
  const CafeDeck* deck() const { return cafe_deck_get(h); }
};
}  // namespace cafe_facade

template <typename T>
class SinglePhaseBase {
 private:
  friend class MultiPhaseDDP<T>;

 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  SinglePhaseBase() {}
  virtual ~SinglePhaseBase() {}
  virtual void initialization() = 0;
  virtual size_t get_state_dim() { return 0; }
  virtual size_t get_control_dim() { return 0; }
  virtual void update_SS_config(int ss_sz) { (void)ss_sz; }
  virtual void empty_control() {}
  virtual void get_trajectory(std::vector<std::vector<float>>& x_tau, std::vector<std::vector<float>>& u_tau) { (void)x_tau; (void)u_tau; }
  virtual void print() {}

  // ---- binding to the GPU path
  std::shared_ptr<cafe_facade::DeckOwner> cafe_deck;   // set by the problem builder that created the phase
  int cafe_phase_index = -1;
  virtual long cafe_record_size() const = 0;                    // doubles this phase takes in the packed solution record
  virtual void cafe_pack_guess(double* rec) const = 0;          // Xbar, Ubar, K -> record (the other arrays zero)
  virtual void cafe_unpack_solution(const double* rec) = 0;     // record -> Xbar, X, Ubar, U, Y, dU, K, Qu, Quu, Qux, G
  // (sigma, lambda) of the up to four elements of this phase's touchdown constraint as the last solve left them: in the reference this state
  // lives in the phase's TouchDownConstraint object and survives every MPC update (reset_params is empty, ConstraintsBase.h:367-374);
  // <Problem>::update() hands it to the phase that continues this one, MultiPhaseDDP::solve starts from it (cafe_gpu_set_al_params)
  double cafe_al[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool cafe_al_set = false;
};
