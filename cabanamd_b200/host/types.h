// types.h — scalar types and option enums of the host layer.
// Mirrors the names of the reference's src/types.h:54-150 so code written against
// CabanaMD's headers compiles against these; FP64 is the default build
// (T_FLOAT = double), matching types.h:133-148.
#ifndef CBMD_HOST_TYPES_H
#define CBMD_HOST_TYPES_H

typedef int T_INT;
#ifdef CBMD_SINGLE_PRECISION
#error "the FP32 build variant is not available yet; libcbmd_cuda computes in FP64"
#endif
typedef double T_FLOAT;
typedef T_FLOAT T_X_FLOAT;
typedef T_FLOAT T_V_FLOAT;
typedef T_FLOAT T_F_FLOAT;

// run-time options (values are only compared by name, never serialised)
enum { FORCE_LJ, FORCE_SNAP, FORCE_NNP };
enum { FORCE_ITER_NEIGH_FULL, FORCE_ITER_NEIGH_HALF };
enum { FORCE_PARALLEL_NEIGH_SERIAL, FORCE_PARALLEL_NEIGH_TEAM, FORCE_PARALLEL_NEIGH_VECTOR };
enum { NEIGH_NONE, NEIGH_VERLET_2D, NEIGH_VERLET_CSR, NEIGH_TREE_2D, NEIGH_TREE_CSR };
enum { INPUT_LAMMPS };
enum { UNITS_REAL, UNITS_LJ, UNITS_METAL };
enum { LATTICE_SC, LATTICE_FCC };
enum { INTEGRATOR_NVE };
enum { BINNING_LINKEDCELL };
enum { COMM_MPI };
enum { DEFAULT, SERIAL, PTHREAD, OPENMP, CUDA, HIP };

#endif
