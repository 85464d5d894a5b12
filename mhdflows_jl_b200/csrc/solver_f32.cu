// Float32 instantiation of the solver (T = Float32 is the reference default, pgen.jl:90).
#include "solver.cuh"

mhdf_handle* mhdf_make_solver_f32(const mhdf_config& c) { return new Solver<float>(c); }
