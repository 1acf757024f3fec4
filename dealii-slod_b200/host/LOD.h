// Host side of the B200-native SLOD offline phase: a deal.II-free mirror of the reference's driver classes
//
//   LODParameters<dim,spacedim>   include/LOD.h:85-157    same members, same .prm keys (section "Problem")
//   Patch<dim>                    include/LOD.h:68-82
//   LOD<dim,spacedim>             include/LOD.h:159-262   run() = the same stage sequence (source/LOD.cc:1423-1441)
//   problem_parameter<dim>        include/Diffusion.h:7-54
//   DiffusionProblem              include/Diffusion.h:56-306
//   ElasticityProblem             include/Elasticity.h:92-438
//
// The two hot members, compute_basis_function_candidates() (source/LOD.cc:296-768) and assemble_global_matrix()
// (source/LOD.cc:860-973), are calls into the C ABI of libslod_b200.so (include/slod.h); everything the reference
// builds per patch on the host (Triangulation, DoFHandler, sparsity patterns, index vectors) is closed-form index
// arithmetic inside that library, so the host keeps only the parameter interface, the coefficient tables and the
// results.  After the offline phase run() also does SURVEY.md section 8f row 1 through the same handle: fine right-hand
// side (constant forcing), solve() = C^T f + coarse CG, the prolongation C u and -- with "Compare with fine global
// solution" -- section 8f row 2: the fine FEM reference solve and the SLOD-vs-FEM(h) error norms.  The other error
// tables (exact solution, coarse FEM) and the VTU writers are out of scope and are not run.
//
// Error behaviour: like the reference (AssertThrow -> exception -> main prints and returns 1), every failing C ABI
// call becomes a std::runtime_error carrying slod_last_error().
#pragma once

#include <slod.h>

#include "vtu.h"

#include <chrono>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace slodhost {

// ----------------------------------------------------------------------------------------------------------------
// A minimal ParameterHandler: "subsection X" / "end" / "set Key = value" / '#' comments, as written by deal.II's
// ParameterAcceptor (the reference ships no .prm; ParameterAcceptor::initialize writes a template when the file is
// missing, README:3 -- so does this one).
// ----------------------------------------------------------------------------------------------------------------
class ParameterHandler {
public:
  void declare(const std::string &path, const std::string &key, const std::string &def, const std::string &doc = "") {
    const std::string full = path + "/" + key;
    if (!values.count(full)) order.push_back(full);
    values[full] = def;
    docs[full] = doc;
  }
  std::string get(const std::string &path, const std::string &key) const {
    auto it = values.find(path + "/" + key);
    if (it == values.end()) throw std::runtime_error("undeclared parameter <" + path + "/" + key + ">");
    return it->second;
  }
  long get_integer(const std::string &p, const std::string &k) const {
    const std::string v = get(p, k);
    char *end = nullptr;
    const long r = std::strtol(v.c_str(), &end, 10);
    if (end == v.c_str() || *end != '\0')
      throw std::runtime_error("parameter <" + k + "> = '" + v + "' is not an integer");
    return r;
  }
  double get_double(const std::string &p, const std::string &k) const {
    const std::string v = get(p, k);
    char *end = nullptr;
    const double r = std::strtod(v.c_str(), &end);
    if (end == v.c_str() || *end != '\0') throw std::runtime_error("parameter <" + k + "> = '" + v + "' is not a number");
    return r;
  }
  bool get_bool(const std::string &p, const std::string &k) const {
    const std::string v = get(p, k);
    if (v == "true" || v == "yes" || v == "on" || v == "1") return true;
    if (v == "false" || v == "no" || v == "off" || v == "0") return false;
    throw std::runtime_error("parameter <" + k + "> = '" + v + "' is not a boolean");
  }
  // returns false if the file does not exist
  bool parse_file(const std::string &file) {
    std::ifstream in(file);
    if (!in) return false;
    std::vector<std::string> stack;
    std::string line;
    int lineno = 0;
    while (std::getline(in, line)) {
      ++lineno;
      const auto hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      line = trim(line);
      if (line.empty()) continue;
      if (line.rfind("subsection", 0) == 0) {
        stack.push_back(trim(line.substr(10)));
      } else if (line == "end") {
        if (stack.empty()) throw std::runtime_error(file + ":" + std::to_string(lineno) + ": 'end' without subsection");
        stack.pop_back();
      } else if (line.rfind("set", 0) == 0) {
        const auto eq = line.find('=');
        if (eq == std::string::npos) throw std::runtime_error(file + ":" + std::to_string(lineno) + ": missing '='");
        const std::string key = trim(line.substr(3, eq - 3)), val = trim(line.substr(eq + 1));
        std::string path;
        for (const auto &s : stack) path += "/" + s;
        const std::string full = path + "/" + key;
        if (!values.count(full))
          throw std::runtime_error(file + ":" + std::to_string(lineno) + ": no such parameter <" + full + ">");
        values[full] = val;
      } else {
        throw std::runtime_error(file + ":" + std::to_string(lineno) + ": cannot parse '" + line + "'");
      }
    }
    if (!stack.empty()) throw std::runtime_error(file + ": unterminated subsection '" + stack.back() + "'");
    return true;
  }
  // ParameterHandler::Short style output
  void print_parameters(const std::string &file) const {
    std::ofstream out(file);
    if (!out) throw std::runtime_error("cannot write " + file);
    std::vector<std::string> open;
    for (const auto &full : order) {
      std::vector<std::string> parts;
      std::stringstream ss(full.substr(1));
      std::string item;
      while (std::getline(ss, item, '/')) parts.push_back(item);
      const std::string key = parts.back();
      parts.pop_back();
      size_t common = 0;
      while (common < open.size() && common < parts.size() && open[common] == parts[common]) ++common;
      while (open.size() > common) {
        open.pop_back();
        out << std::string(2 * open.size(), ' ') << "end\n";
      }
      while (open.size() < parts.size()) {
        out << std::string(2 * open.size(), ' ') << "subsection " << parts[open.size()] << "\n";
        open.push_back(parts[open.size()]);
      }
      out << std::string(2 * open.size(), ' ') << "set " << key << " = " << values.at(full) << "\n";
    }
    while (!open.empty()) {
      open.pop_back();
      out << std::string(2 * open.size(), ' ') << "end\n";
    }
  }

private:
  static std::string trim(const std::string &s) {
    const auto a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
  }
  std::map<std::string, std::string> values, docs;
  std::vector<std::string> order;
};

// ----------------------------------------------------------------------------------------------------------------
// LODParameters  (include/LOD.h:85-157).  Keys marked [+] are the ones the reference has commented out
// (include/LOD.h:100-102, 144-147; README:14) and hard-codes in the problem constructors; here they are real
// parameters whose defaults are the hard-coded values.
// ----------------------------------------------------------------------------------------------------------------
template <int dim, int spacedim>
class LODParameters {
public:
  LODParameters() {
    const std::string P = "/Problem";
    prm.declare(P, "Output directory", ".");
    prm.declare(P, "Output name", "solution");
    prm.declare(P, "Oversampling", "1");
    prm.declare(P, "Number of subdivisions", "2");
    prm.declare(P, "Number of global refinements", "2");
    prm.declare(P, "Compare with fine global solution", "false");
    prm.declare(P, "Stabilize phi_LOD candidates", "false");
    prm.declare(P + "/Coefficients", "Constant problem coefficients", "true");
    prm.declare(P + "/Coefficients", "Minimum value for random coefficients", "1");                          // [+]
    prm.declare(P + "/Coefficients", "Maximum value for random coefficients", "100");                        // [+]
    prm.declare(P + "/Coefficients", "Refinement for random coefficients", spacedim == 1 ? (dim == 2 ? "8" : "6") : "6");  // [+]
    prm.declare(P + "/Coefficients", "Random seed", "0");  // [+] 0: unseeded rand() like the reference
    // ParsedFunction (include/LOD.h:104,123): components separated by ';'.  This host evaluates constants only.
    prm.declare(P + "/Right hand side", "Function expression", spacedim == 1 ? "0" : "0; 0");
    // ReductionControl (include/LOD.h:109,127) with deal.II's defaults
    prm.declare(P + "/Solver/Coarse solver control", "Max steps", "100");
    prm.declare(P + "/Solver/Coarse solver control", "Tolerance", "1e-10");
    prm.declare(P + "/Solver/Coarse solver control", "Reduction", "1e-2");
    prm.declare(P + "/Solver/Fine solver control", "Max steps", "100");      // include/LOD.h:108,126
    prm.declare(P + "/Solver/Fine solver control", "Tolerance", "1e-10");
    prm.declare(P + "/Solver/Fine solver control", "Reduction", "1e-2");
    prm.declare(P + "/B200", "Device", "-1");              // [+] CUDA device ordinal, -1 = current
    prm.declare(P + "/B200", "Number of GPUs", "1");       // [+] devices driven by the one handle (NCCL inside the library)
    prm.declare(P + "/B200", "Write coarse matrix", "true");
    prm.declare(P + "/B200", "Write VTU output", "true");  // [+] the reference always writes its three .vtu files
  }
  // ParameterAcceptor::initialize(prm_file): parse, or write a template and keep the defaults
  void initialize(const std::string &prm_file) {
    if (!prm.parse_file(prm_file)) prm.print_parameters(prm_file);
    const std::string P = "/Problem";
    output_directory = prm.get(P, "Output directory");
    output_name = prm.get(P, "Output name");
    oversampling = (unsigned)nonneg(prm.get_integer(P, "Oversampling"), "Oversampling");
    n_subdivisions = (unsigned)nonneg(prm.get_integer(P, "Number of subdivisions"), "Number of subdivisions");
    n_global_refinements =
        (unsigned)nonneg(prm.get_integer(P, "Number of global refinements"), "Number of global refinements");
    solve_fine_problem = prm.get_bool(P, "Compare with fine global solution");
    LOD_stabilization = prm.get_bool(P, "Stabilize phi_LOD candidates");
    constant_coefficients = prm.get_bool(P + "/Coefficients", "Constant problem coefficients");
    random_value_min = prm.get_double(P + "/Coefficients", "Minimum value for random coefficients");
    random_value_max = prm.get_double(P + "/Coefficients", "Maximum value for random coefficients");
    random_value_refinement =
        (unsigned)nonneg(prm.get_integer(P + "/Coefficients", "Refinement for random coefficients"), "Refinement");
    random_seed = (unsigned)nonneg(prm.get_integer(P + "/Coefficients", "Random seed"), "Random seed");
    rhs_expression = prm.get(P + "/Right hand side", "Function expression");
    coarse_max_steps = (unsigned)nonneg(prm.get_integer(P + "/Solver/Coarse solver control", "Max steps"), "Max steps");
    coarse_tolerance = prm.get_double(P + "/Solver/Coarse solver control", "Tolerance");
    coarse_reduction = prm.get_double(P + "/Solver/Coarse solver control", "Reduction");
    fine_max_steps = (unsigned)nonneg(prm.get_integer(P + "/Solver/Fine solver control", "Max steps"), "Max steps");
    fine_tolerance = prm.get_double(P + "/Solver/Fine solver control", "Tolerance");
    fine_reduction = prm.get_double(P + "/Solver/Fine solver control", "Reduction");
    device = (int)prm.get_integer(P + "/B200", "Device");
    n_gpus = (int)nonneg(prm.get_integer(P + "/B200", "Number of GPUs"), "Number of GPUs");
    write_coarse_matrix = prm.get_bool(P + "/B200", "Write coarse matrix");
    write_vtu = prm.get_bool(P + "/B200", "Write VTU output");
    (void)rhs_constants();   // refuse what this host cannot evaluate before any work is done
  }

  std::string output_directory = ".";
  std::string output_name = "solution";
  unsigned int oversampling = 1;
  unsigned int n_subdivisions = 2;
  unsigned int n_global_refinements = 2;
  bool solve_fine_problem = false;
  bool LOD_stabilization = false;
  bool constant_coefficients = true;
  double random_value_min = 1;
  double random_value_max = 100;
  unsigned int random_value_refinement = 8;
  unsigned int random_seed = 0;
  std::string rhs_expression = spacedim == 1 ? "0" : "0; 0";
  unsigned int coarse_max_steps = 100;
  double coarse_tolerance = 1e-10, coarse_reduction = 1e-2;
  unsigned int fine_max_steps = 100;
  double fine_tolerance = 1e-10, fine_reduction = 1e-2;
  int device = -1;
  int n_gpus = 1;
  bool write_coarse_matrix = true;
  bool write_vtu = true;

  // the constant value of every component of the right-hand side; anything but numbers is refused
  std::vector<double> rhs_constants() const {
    std::vector<double> v;
    std::stringstream ss(rhs_expression);
    std::string item;
    while (std::getline(ss, item, ';')) {
      size_t used = 0;
      double x = 0;
      try {
        x = std::stod(item, &used);
      } catch (const std::exception &) {
        used = 0;
      }
      while (used < item.size() && std::isspace((unsigned char)item[used])) ++used;
      if (used == 0 || used != item.size())
        throw std::runtime_error("parameter <Function expression> of the right-hand side: only constant expressions are "
                                 "evaluated by this host, got '" + item + "'");
      v.push_back(x);
    }
    if ((int)v.size() != spacedim)
      throw std::runtime_error("parameter <Function expression> of the right-hand side needs " +
                               std::to_string(spacedim) + " components");
    return v;
  }

  mutable ParameterHandler prm;

private:
  static long nonneg(long v, const char *name) {
    if (v < 0) throw std::runtime_error(std::string("parameter <") + name + "> must not be negative");
    return v;
  }
};

// include/LOD.h:68-82.  `cells` holds active-cell indices (Morton order) instead of cell iterators; the vectors are
// in deal.II's patch-local DoF numbering (dh_fine_patch.distribute_dofs, source/LOD.cc:365-366).
template <int dim>
class Patch {
public:
  std::vector<unsigned int> cells;
  std::vector<std::vector<double>> basis_function;
  std::vector<std::vector<double>> basis_function_premultiplied;
  unsigned int contained_patches = 0;
};

// include/Diffusion.h:7-54: piecewise-constant random field on a 2^r grid, values drawn with libc rand()
template <int dim>
class problem_parameter {
public:
  problem_parameter(double min, double max, unsigned int r) : min_val(min), max_val(max), refinement(r) {
    N_cells_per_line = 1u << refinement;
    eta = 1.0 / N_cells_per_line;
    size_t N_cells = 1;
    for (int a = 0; a < dim; ++a) N_cells *= N_cells_per_line;
    if (max_val != min_val) {
      random_values.reserve(N_cells);
      for (size_t i = 0; i < N_cells; ++i) {
        const double v = min_val + static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (max_val - min_val)));
        random_values.push_back(v);
      }
    } else {
      random_values.assign(N_cells, min_val);
    }
  }
  // value at a point (include/Diffusion.h:40-53), dim-generic index x + N y (+ N^2 z)
  double value(const double *p) const {
    if (max_val == min_val) return min_val;
    size_t idx = 0, mul = 1;
    for (int a = 0; a < dim; ++a) {
      idx += (size_t)std::floor(p[a] / eta) * mul;
      mul *= N_cells_per_line;
    }
    return random_values[idx];
  }
  const std::vector<double> &table() const { return random_values; }
  unsigned int get_refinement() const { return refinement; }

private:
  const double min_val, max_val;
  const unsigned int refinement;
  std::vector<double> random_values;
  unsigned int N_cells_per_line;
  double eta;
};

// TimerOutput(summary, wall_times) stand-in with the reference's section names (source/LOD.cc:126,301,864)
class TimerOutput {
public:
  class Scope {
  public:
    Scope(TimerOutput &t, const std::string &name) : t(t), name(name), t0(std::chrono::steady_clock::now()) {}
    ~Scope() {
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (!t.sections.count(name)) t.order.push_back(name);
      t.sections[name] += s;
    }

  private:
    TimerOutput &t;
    std::string name;
    std::chrono::steady_clock::time_point t0;
  };
  void print_summary(std::ostream &out) const {
    out << "+---------------------------------------------+------------+\n";
    out << "| Section                                     | wall time  |\n";
    out << "+---------------------------------------------+------------+\n";
    for (const auto &n : order) {
      char buf[128];
      std::snprintf(buf, sizeof buf, "| %-43s | %9.4fs |\n", n.c_str(), sections.at(n));
      out << buf;
    }
    out << "+---------------------------------------------+------------+\n";
  }
  std::map<std::string, double> sections;
  std::vector<std::string> order;
};

struct CoarseMatrix {  // global_stiffness_matrix (include/LOD.h:234) as CSR
  std::vector<int64_t> rowptr, col;
  std::vector<double> val;
  int64_t n_rows = 0;
  double frobenius_norm() const {
    double s = 0;
    for (double v : val) s += v * v;
    return std::sqrt(s);
  }
};

// ----------------------------------------------------------------------------------------------------------------
template <int dim, int spacedim>
class LOD {
public:
  explicit LOD(const LODParameters<dim, spacedim> &par) : par(par), pcout(std::cout) {}
  virtual ~LOD() {
    if (slod) slod_destroy(slod);
  }

  virtual void run() {
    print_parameters();
    make_grid();
    make_fe();
    initialize_patches();
    create_random_problem_coefficients();
    output_coefficients();
    compute_basis_function_candidates();
    assemble_global_matrix();
    assemble_and_solve_fem_problem();
    solve();
    compare_lod_with_fem();
    output_coarse_results();
    output_fine_results();
    output_offline_results();
    if (par.solve_fine_problem) {   // source/LOD.cc:1463-1465
      pcout << "SLOD vs reference FEM(h)" << std::endl;
      char buf[200];
      std::snprintf(buf, sizeof buf,
                    "cells dofs   u_L2_norm    u_Linfty_norm u_H1_norm    u_energy_norm\n%5lld %6zu %12.6e %12.6e %12.6e %12.6e",
                    (long long)n_patches, n_dofs_fine, error_LOD_FEMh[0], error_LOD_FEMh[1], error_LOD_FEMh[2],
                    error_LOD_FEMh[3]);
      pcout << buf << std::endl;
    }
    computing_timer.print_summary(pcout);
  }

  const CoarseMatrix &get_global_stiffness_matrix() const { return global_stiffness_matrix; }
  const std::vector<Patch<dim>> &get_patches() const { return patches; }
  const std::vector<double> &get_solution() const { return solution; }
  const std::vector<double> &get_lod_solution() const { return lod_solution; }

protected:
  void check(int rc, const char *what) const {
    if (rc != SLOD_OK)
      throw std::runtime_error(std::string(what) + ": " + (slod ? slod_last_error(slod) : slod_last_create_error()));
  }

  void print_parameters() const {  // source/LOD.cc:33-63
    if (spacedim == 1)
      pcout << "Running LOD Diffusion problem in " << dim << "D" << std::endl;
    else
      pcout << "Running LOD Elasticity problem in " << dim << "D" << std::endl;
    par.prm.print_parameters(par.output_directory + "/" + "used_parameters_" + std::to_string(dim) + ".prm");
  }

  void make_grid() {  // source/LOD.cc:108-119: unit hyper-cube, refine_global -> the handle
    slod_params p{};
    p.dim = dim;
    p.spacedim = spacedim;
    p.n_global_refinements = (int)par.n_global_refinements;
    p.n_subdivisions = (int)par.n_subdivisions;
    p.oversampling = (int)par.oversampling;
    p.stabilize = par.LOD_stabilization ? 1 : 0;
    p.problem = (spacedim == 1) ? SLOD_PROBLEM_DIFFUSION : SLOD_PROBLEM_ELASTICITY;
    p.quirk_presaved = par.constant_coefficients ? 1 : 0;  // source/LOD.cc:354-362
    p.device = par.device;
    p.n_gpus = par.n_gpus;   // locally_owned_patches over the devices of this process (source/LOD.cc:116-118)
    check(slod_create(&p, &slod), "slod_create");
    int64_t n = 0;
    check(slod_patch_count(slod, &n), "slod_patch_count");
    n_patches = n;
    pcout << "Number of coarse cell = " << n_patches << std::endl;
  }

  void make_fe() {  // source/LOD.cc:65-106: sizes only -- FE_DGQ(0)^s coarse, FE_Q_iso_Q1(n)^s fine
    size_t fine_nodes = 1;
    for (int a = 0; a < dim; ++a) fine_nodes *= ((size_t)par.n_subdivisions << par.n_global_refinements) + 1;
    n_dofs_coarse = (size_t)spacedim * n_patches;
    n_dofs_fine = (size_t)spacedim * fine_nodes;
    pcout << "Number of coarse dofs = " << n_dofs_coarse << ", fine dofs = " << n_dofs_fine << std::endl;
  }

  void create_patches() {  // source/LOD.cc:122-244
    TimerOutput::Scope t(computing_timer, "1: Create Patches");
    patches.resize(n_patches);
    size_t size_biggest_patch = 0, size_tiniest_patch = (size_t)-1;
    for (int64_t id = 0; id < n_patches; ++id) {
      int32_t nc = 0;
      check(slod_get_patch_cells(slod, id, nullptr, &nc), "slod_get_patch_cells");
      std::vector<uint32_t> cells(nc);
      check(slod_get_patch_cells(slod, id, cells.data(), &nc), "slod_get_patch_cells");
      patches[id].cells.assign(cells.begin(), cells.end());
      size_biggest_patch = std::max<size_t>(size_biggest_patch, nc);
      size_tiniest_patch = std::min<size_t>(size_tiniest_patch, nc);
    }
    pcout << "Number of patches = " << patches.size() << ", patch sizes (" << size_tiniest_patch << ", "
          << size_biggest_patch << ")" << std::endl;
  }

  void initialize_patches() {  // source/LOD.cc:1380-1393 (create_mesh_for_patch is implicit index arithmetic)
    create_patches();
    for (auto &p : patches) {
      p.basis_function.assign(spacedim, {});
      p.basis_function_premultiplied.assign(spacedim, {});
    }
  }

  // hooks of the problem classes: the coefficient tables replace the virtual assemble_stiffness
  // (include/LOD.h:205-211): the PDE enters through them
  virtual void create_random_problem_coefficients() {}
  virtual unsigned int n_coefficient_fields() const = 0;
  virtual const problem_parameter<dim> &coefficient(unsigned int field) const = 0;

  void compute_basis_function_candidates() {  // source/LOD.cc:296-768
    TimerOutput::Scope t(computing_timer, "2: compute basis function (B200)");
    for (unsigned int f = 0; f < n_coefficient_fields(); ++f) {
      const auto &c = coefficient(f);
      check(slod_set_coefficient(slod, (int)f, (int)c.get_refinement(), c.table().data(), c.table().size()),
            "slod_set_coefficient");
    }
    check(slod_compute_basis(slod), "slod_compute_basis");
    std::vector<double> phi, aphi;
    std::vector<uint32_t> loc;
    for (int64_t id = 0; id < n_patches; ++id) {
      int32_t nf = 0;
      check(slod_get_patch_local_dofs(slod, id, nullptr, &nf), "slod_get_patch_local_dofs");
      loc.resize(nf);
      phi.resize(nf);
      aphi.resize(nf);
      check(slod_get_patch_local_dofs(slod, id, loc.data(), &nf), "slod_get_patch_local_dofs");
      for (int d = 0; d < spacedim; ++d) {
        check(slod_get_basis(slod, id, d, phi.data(), aphi.data()), "slod_get_basis");
        auto &bf = patches[id].basis_function[d];
        auto &bp = patches[id].basis_function_premultiplied[d];
        bf.assign(nf, 0.0);
        bp.assign(nf, 0.0);
        for (int i = 0; i < nf; ++i) {
          bf[loc[i]] = phi[i];
          bp[loc[i]] = aphi[i];
        }
      }
    }
  }

  void assemble_global_matrix() {  // source/LOD.cc:860-973
    TimerOutput::Scope t(computing_timer, "3: Assemble global matrix (B200)");
    check(slod_assemble_coarse(slod), "slod_assemble_coarse");
    int64_t n_rows = 0, nnz = 0;
    check(slod_get_coarse_csr(slod, nullptr, nullptr, nullptr, &n_rows, &nnz), "slod_get_coarse_csr");
    auto &K = global_stiffness_matrix;
    K.n_rows = n_rows;
    K.rowptr.resize(n_rows + 1);
    K.col.resize(nnz);
    K.val.resize(nnz);
    check(slod_get_coarse_csr(slod, K.rowptr.data(), K.col.data(), K.val.data(), &n_rows, &nnz),
          "slod_get_coarse_csr");
  }

  // The right-hand side half of assemble_and_solve_fem_problem (source/LOD.cc:1004-1040): F_i = int f phi_i with zero
  // rows on the domain boundary.  For a constant f the Gauss sums (include/Diffusion.h:188-191) are the tensor product
  // of the 1-D weights h (interior node) -- evaluated here on the host, it is mesh set-up, not patch work.
  void assemble_fem_rhs() {
    const std::vector<double> f = par.rhs_constants();
    const size_t G = ((size_t)par.n_subdivisions << par.n_global_refinements) + 1;
    const double h = 1.0 / (double)(G - 1);
    fem_rhs.assign(n_dofs_fine, 0.0);
    size_t nodes = 1;
    for (int a = 0; a < dim; ++a) nodes *= G;
    for (size_t node = 0; node < nodes; ++node) {
      size_t r = node;
      double w = 1.0;
      for (int a = 0; a < dim; ++a) {
        const size_t i = r % G;
        r /= G;
        w *= (i == 0 || i == G - 1) ? 0.0 : h;
      }
      for (int c = 0; c < spacedim; ++c) fem_rhs[node * spacedim + c] = w * f[c];
    }
    double nrm = 0;
    for (double v : fem_rhs) nrm += v * v;
    pcout << "     fem rhs l2 norm = " << std::sqrt(nrm) << std::endl;   // source/LOD.cc:1039
  }

  // source/LOD.cc:1004-1094: right-hand side always (solve() needs it), the fine solve only when
  // "Compare with fine global solution" is set (source/LOD.cc:1043)
  void assemble_and_solve_fem_problem() {
    assemble_fem_rhs();
    if (!par.solve_fine_problem) return;
    TimerOutput::Scope t(computing_timer, "0: fine FEM solve (B200)");
    fem_solution.assign(n_dofs_fine, 0.0);
    int32_t steps = 0;
    double residual = 0;
    check(slod_fem_solve(slod, fem_rhs.data(), fem_solution.data(), (int32_t)par.fine_max_steps, par.fine_tolerance,
                         par.fine_reduction, &steps, &residual),
          "slod_fem_solve");
    pcout << "   size of fem u " << fem_solution.size() << std::endl;   // source/LOD.cc:1092
    char buf[128];
    std::snprintf(buf, sizeof buf, "   fine CG: %d steps, residual %.3e", (int)steps, residual);
    pcout << buf << std::endl;
  }

  void compare_lod_with_fem() {  // source/LOD.cc:1240-1260
    prolongate_lod_solution();
    if (!par.solve_fine_problem) return;
    TimerOutput::Scope t(computing_timer, "5: compare FEM vs LOD (B200)");
    std::vector<double> diff(n_dofs_fine);
    for (size_t i = 0; i < n_dofs_fine; ++i) diff[i] = fem_solution[i] - lod_solution[i];
    // par.error_LOD_FEMh.difference(dh, fem_solution, lod_solution) (source/LOD.cc:1252): the table's default norms
    // L2, Linfty, H1 with the reference's quadrature; [+] the energy norm of the problem's own bilinear form (exact)
    double l2 = 0, linf = 0, h1 = 0, en = 0;
    check(slod_fine_norms_reference(slod, diff.data(), &l2, &linf, &h1), "slod_fine_norms_reference");
    check(slod_fine_norms(slod, diff.data(), nullptr, nullptr, &en), "slod_fine_norms");
    error_LOD_FEMh[0] = l2;
    error_LOD_FEMh[1] = linf;
    error_LOD_FEMh[2] = h1;
    error_LOD_FEMh[3] = en;
  }

  void solve() {  // source/LOD.cc:975-1001
    TimerOutput::Scope t(computing_timer, "4: Solve LOD (B200)");
    system_rhs.assign(n_dofs_coarse, 0.0);
    solution.assign(n_dofs_coarse, 0.0);
    check(slod_coarse_rhs(slod, fem_rhs.data(), system_rhs.data()), "slod_coarse_rhs");
    double nrm = 0;
    for (double v : system_rhs) nrm += v * v;
    pcout << "     rhs l2 norm = " << std::sqrt(nrm) << std::endl;
    int32_t steps = 0;
    double residual = 0;
    check(slod_coarse_solve(slod, system_rhs.data(), solution.data(), (int32_t)par.coarse_max_steps,
                            par.coarse_tolerance, par.coarse_reduction, &steps, &residual),
          "slod_coarse_solve");
    pcout << "   size of u " << solution.size() << std::endl;
    char buf[128];
    std::snprintf(buf, sizeof buf, "   coarse CG: %d steps, residual %.3e", (int)steps, residual);
    pcout << buf << std::endl;
  }

  void prolongate_lod_solution() {  // first lines of compare_lod_with_fem, source/LOD.cc:1247-1251
    lod_solution.assign(n_dofs_fine, 0.0);
    check(slod_prolongate(slod, solution.data(), lod_solution.data()), "slod_prolongate");
    double nrm = 0;
    for (double v : lod_solution) nrm += v * v;
    char buf[128];
    std::snprintf(buf, sizeof buf, "   lod solution l2 norm = %.12e", std::sqrt(nrm));
    pcout << buf << std::endl;
  }

  // <name>_coefficients.vtu (include/Diffusion.h:70-108): the coefficient fields as cell data on the fine sub-cell mesh
  // of (2^ref n)^dim cells, sampled at the cell centres like VectorTools::interpolate on FE_DGQ(0) does
  void output_coefficients() {
    if (!par.write_vtu) return;
    const int nc = (int)((1u << par.n_global_refinements) * par.n_subdivisions);
    const size_t ncell = (size_t)std::pow((double)nc, dim);
    const char *names[2] = {spacedim == 1 ? "alpha" : "lambda", "mu"};
    std::vector<std::vector<double>> vals(n_coefficient_fields(), std::vector<double>(ncell));
    for (unsigned f = 0; f < n_coefficient_fields(); ++f)
      for (size_t c = 0; c < ncell; ++c) {
        size_t r = c;
        double p[3] = {0, 0, 0};
        for (int a = 0; a < dim; ++a) {
          p[a] = ((double)(r % nc) + 0.5) / nc;
          r /= nc;
        }
        vals[f][c] = coefficient(f).value(p);
      }
    std::vector<vtu::Field> cell;
    for (unsigned f = 0; f < n_coefficient_fields(); ++f) cell.push_back({names[f], 1, vals[f].data()});
    vtu::write(par.output_directory + "/" + par.output_name + "_coefficients.vtu", dim, nc, {}, cell);
  }

  // <name>_coarse.vtu (source/LOD.cc:248-293): the coarse solution, one value per cell and component (FE_DGQ(0)^s),
  // and the exact solution interpolated the same way (constant expressions only on this host)
  void output_coarse_results() {
    if (!par.write_vtu) return;
    const int N = 1 << par.n_global_refinements;
    const size_t ncell = (size_t)n_patches;
    std::vector<double> sol(ncell * spacedim), exact(ncell * spacedim, 0.0);
    for (size_t c = 0; c < ncell; ++c) {   // lexicographic cell -> patch id (Morton)
      size_t r = c, code = 0;
      int idx[3] = {0, 0, 0};
      for (int a = 0; a < dim; ++a) {
        idx[a] = (int)(r % N);
        r /= N;
      }
      for (unsigned b = 0; b < par.n_global_refinements; ++b)
        for (int a = 0; a < dim; ++a) code |= (size_t)((idx[a] >> b) & 1) << (dim * b + a);
      for (int d = 0; d < spacedim; ++d) sol[c * spacedim + d] = solution[code * spacedim + d];
    }
    vtu::write(par.output_directory + "/" + par.output_name + "_coarse.vtu", dim, N, {},
               {{"LOD_solution", spacedim, sol.data()}, {"exact_solution", spacedim, exact.data()}});
  }

  // <name>_fine.vtu (source/LOD.cc:1262-1377): fem_reference, exact_rhs and lod_solution as point data on the fine
  // grid.  Every sub-cell is written (the reference's build_patches() keeps the coarse-cell vertices only); the
  // coarse FEM solution of the reference's table is not computed by this host.
  void output_fine_results() {
    if (!par.write_vtu) return;
    const int nc = (int)((1u << par.n_global_refinements) * par.n_subdivisions);
    std::vector<double> rhs_nodal(n_dofs_fine);
    const std::vector<double> fc = par.rhs_constants();
    for (size_t i = 0; i < n_dofs_fine; ++i) rhs_nodal[i] = fc[i % spacedim];
    std::vector<vtu::Field> pt;
    if (par.solve_fine_problem) pt.push_back({"fem_reference", spacedim, fem_solution.data()});
    pt.push_back({"exact_rhs", spacedim, rhs_nodal.data()});
    pt.push_back({"lod_solution", spacedim, lod_solution.data()});
    vtu::write(par.output_directory + "/" + par.output_name + "_fine.vtu", dim, nc, pt, {});
  }

  void output_offline_results() {
    const auto &K = global_stiffness_matrix;
    char buf[256];
    std::snprintf(buf, sizeof buf, "global_stiffness_matrix: %lld x %lld, %lld nonzeros, frobenius norm = %.12e",
                  (long long)K.n_rows, (long long)K.n_rows, (long long)K.val.size(), K.frobenius_norm());
    pcout << buf << std::endl;
    double ms[8] = {0};
    slod_get_timings(slod, ms, 8);
    std::snprintf(buf, sizeof buf, "device ms: solve %.3f dense %.3f select %.3f finish %.3f coarse %.3f", ms[0], ms[1],
                  ms[2], ms[3], ms[4]);
    pcout << buf << std::endl;
    if (par.write_coarse_matrix) {
      const std::string file = par.output_directory + "/" + par.output_name + "_coarse_matrix.bin";
      std::ofstream out(file, std::ios::binary);
      if (!out) throw std::runtime_error("cannot write " + file);
      const int64_t hdr[2] = {K.n_rows, (int64_t)K.val.size()};
      out.write((const char *)hdr, sizeof hdr);
      out.write((const char *)K.rowptr.data(), sizeof(int64_t) * K.rowptr.size());
      out.write((const char *)K.col.data(), sizeof(int64_t) * K.col.size());
      out.write((const char *)K.val.data(), sizeof(double) * K.val.size());
    }
  }

  const LODParameters<dim, spacedim> &par;
  std::ostream &pcout;
  mutable TimerOutput computing_timer;
  slod_ctx *slod = nullptr;
  int64_t n_patches = 0;
  size_t n_dofs_coarse = 0, n_dofs_fine = 0;
  std::vector<Patch<dim>> patches;
  CoarseMatrix global_stiffness_matrix;
  // include/LOD.h:236-239, lexicographic fine numbering
  std::vector<double> fem_rhs, fem_solution, system_rhs, solution, lod_solution;
  double error_LOD_FEMh[4] = {0, 0, 0, 0};   // L2, Linfty, H1 (reference quadrature), energy norm of fem_solution - lod_solution (include/LOD.h:115)
};

// include/Diffusion.h:56-306.  The reference draws Alpha(1,100,8) in the constructor; here the table is drawn in
// create_random_problem_coefficients() (the first rand() calls of the process either way) so that the [+] keys apply.
template <int dim, int spacedim>
class DiffusionProblem : public LOD<dim, spacedim> {
public:
  explicit DiffusionProblem(const LODParameters<dim, spacedim> &par) : LOD<dim, spacedim>(par) {}
  typedef LOD<dim, spacedim> lod;

protected:
  void create_random_problem_coefficients() override {
    if (lod::par.random_seed) srand(lod::par.random_seed);
    Alpha.reset(new problem_parameter<dim>(lod::par.random_value_min, lod::par.random_value_max,
                                           lod::par.random_value_refinement));
  }
  unsigned int n_coefficient_fields() const override { return 1; }
  const problem_parameter<dim> &coefficient(unsigned int) const override { return *Alpha; }
  std::unique_ptr<problem_parameter<dim>> Alpha;
};

// include/Elasticity.h:92-438: Lambda then Mu (construction order matters for the rand() stream, :104-105)
template <int dim, int spacedim = dim>
class ElasticityProblem : public LOD<dim, spacedim> {
public:
  explicit ElasticityProblem(const LODParameters<dim, spacedim> &par) : LOD<dim, spacedim>(par) {}
  typedef LOD<dim, spacedim> lod;

protected:
  void create_random_problem_coefficients() override {
    if (lod::par.random_seed) srand(lod::par.random_seed);
    Lambda.reset(new problem_parameter<dim>(lod::par.random_value_min, lod::par.random_value_max,
                                            lod::par.random_value_refinement));
    Mu.reset(new problem_parameter<dim>(lod::par.random_value_min, lod::par.random_value_max,
                                        lod::par.random_value_refinement));
  }
  unsigned int n_coefficient_fields() const override { return 2; }
  const problem_parameter<dim> &coefficient(unsigned int f) const override { return f == 0 ? *Lambda : *Mu; }
  std::unique_ptr<problem_parameter<dim>> Lambda, Mu;
};

// app/main_*.cc body: argv[1] = .prm file (default parameters.prm), exceptions -> message + exit code 1
template <class Problem, class Params>
int run_main(int argc, char *argv[]) {
  try {
    std::string prm_file = (argc > 1) ? argv[1] : "parameters.prm";
    Params par;
    Problem problem(par);
    par.initialize(prm_file);
    problem.run();
  } catch (std::exception &exc) {
    std::cerr << std::endl
              << std::endl
              << "----------------------------------------------------" << std::endl;
    std::cerr << "Exception on processing: " << std::endl
              << exc.what() << std::endl
              << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  } catch (...) {
    std::cerr << std::endl
              << std::endl
              << "----------------------------------------------------" << std::endl;
    std::cerr << "Unknown exception!" << std::endl
              << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  }
  return 0;
}

}  // namespace slodhost
