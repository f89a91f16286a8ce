// quad_reference.cpp — see quad_reference.h. File format: scripts/Reference_matlab/generate_reference.m:131-164
// of the reference ("dt" header, then per sample a block of `name` / values line pairs, closed by status_dur).
#include "quad_reference.h"
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cafe {

static std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// values are parsed as float, exactly like std::stof (glibc strtof), then widened
static void parse_floats(const std::string& line, double* dst, int n) {
  std::istringstream ls(line);
  std::string w;
  for (int i = 0; i < n && (ls >> w); ++i) dst[i] = (double)std::strtof(w.c_str(), nullptr);
}

void QuadReference::load_top_level_data(const std::string& fname, bool reorder, int k0) {
  std::ifstream f(fname);
  if (!f.is_open()) throw std::runtime_error("cannot open reference file " + fname);
  tp_data.clear();
  QuadAugmentedState s;
  std::string line;
  while (std::getline(f, line)) {
    std::string key = trim(line);
    if (key.empty() || !(std::isalpha((unsigned char)key[0]))) continue;
    std::string val;
    if (!std::getline(f, val)) break;
    if (key == "dt") { tp_dt = std::strtof(trim(val).c_str(), nullptr); continue; }
    if (key == "body_state") { s = QuadAugmentedState(); parse_floats(val, s.body_state, 12); }
    else if (key == "jnt_angle") parse_floats(val, s.qJ, 12);
    else if (key == "jnt_vel") parse_floats(val, s.qJd, 12);
    else if (key == "torque") parse_floats(val, s.torque, 12);
    else if (key == "foot_placements") parse_floats(val, s.foot_placements, 12);
    else if (key == "foot_velocities") parse_floats(val, s.foot_velocities, 12);
    else if (key == "foot_height") parse_floats(val, s.foot_heights, 4);
    else if (key == "grf") parse_floats(val, s.grf, 12);
    else if (key == "contact") { std::istringstream ls(val); for (int i = 0; i < 4; ++i) ls >> s.contact[i]; }
    else if (key == "status_dur") { parse_floats(val, s.status_dur, 4); tp_data.push_back(s); }
  }
  if (tp_data.empty()) throw std::runtime_error("no samples in reference file " + fname);
  // body state file order [eul, pos, omega, vWorld] -> [pos, eul, vWorld, omega] (QuadReference.cpp:358-371)
  for (auto& st : tp_data) {
    double b[12];
    for (int i = 0; i < 3; ++i) { b[i] = st.body_state[3 + i]; b[3 + i] = st.body_state[i]; b[6 + i] = st.body_state[9 + i]; b[9 + i] = st.body_state[6 + i]; }
    for (int i = 0; i < 12; ++i) st.body_state[i] = b[i];
  }
  if (reorder) {  // swap left/right within front and hind pairs; joint velocities are zeroed (QuadReference.cpp:373-407)
    auto swap3 = [](double* v) { for (int i = 0; i < 3; ++i) { std::swap(v[i], v[3 + i]); std::swap(v[6 + i], v[9 + i]); } };
    for (auto& st : tp_data) {
      swap3(st.qJ); swap3(st.foot_placements); swap3(st.foot_velocities); swap3(st.grf); swap3(st.torque);
      for (int i = 0; i < 12; ++i) st.qJd[i] = 0;
      std::swap(st.contact[0], st.contact[1]); std::swap(st.contact[2], st.contact[3]);
      std::swap(st.status_dur[0], st.status_dur[1]); std::swap(st.status_dur[2], st.status_dur[3]);
    }
  }
  if (k0 > 0) {
    if ((size_t)k0 >= tp_data.size()) throw std::runtime_error("start offset beyond reference length");
    tp_data.erase(tp_data.begin(), tp_data.begin() + k0);
  }
}

void QuadReference::initialize(float plan_horizon) {
  dt = tp_dt;
  dur = plan_horizon;
  sz = (int)std::round(plan_horizon / dt) + 1;
  if ((size_t)(sz + 1) > tp_data.size()) throw std::runtime_error("reference shorter than the planning horizon");
  data.assign(tp_data.begin(), tp_data.begin() + sz + 1);  // one extra sample, as the reference copies
}

int QuadReference::index_at_t(float t) const {
  int k = (int)std::floor(t / dt);
  if ((double)(float)(t - (float)k * dt) > 0.5 * (double)dt) k++;
  if (k >= sz) k = sz - 1;
  return k;
}
const QuadAugmentedState* QuadReference::get_a_reference_ptr_at_t(float t) const { return &data[index_at_t(t)]; }
void QuadReference::get_contact_at_t(int contact[4], float t) const { const auto& s = data[index_at_t(t)]; for (int i = 0; i < 4; ++i) contact[i] = s.contact[i]; }
void QuadReference::get_contact_duration_at_t(double d[4], float t) const { const auto& s = data[index_at_t(t)]; for (int i = 0; i < 4; ++i) d[i] = s.status_dur[i]; }

}  // namespace cafe
