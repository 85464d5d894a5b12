// Float64 instantiation of the solver.
#include "solver.cuh"

mhdf_handle* mhdf_make_solver_f64(const mhdf_config& c) { return new Solver<double>(c); }
