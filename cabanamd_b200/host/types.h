// types.h — scalar types and option enums of the host layer.
// Mirrors the names of the reference's src/types.h:54-150 so code written against
// CabanaMD's headers compiles against these; FP64 is the default build
// (T_FLOAT = double), matching types.h:133-148.
#ifndef CBMD_HOST_TYPES_H
#define CBMD_HOST_TYPES_H

typedef int T_INT;
// FP32 build variant (reference src/types.h:133-148, -DT_F_FLOAT=float -DT_X_FLOAT=float): built
// with -DCBMD_SINGLE_PRECISION (make -C cabanamd_b200/host cbnMD_f32).  The pair forces are then
// evaluated in FP32 on float positions (libcbmd_cuda option "precision" 32, k_force_full_f32);
// the integration state on the device and the host-side arrays stay FP64, because the C ABI
// exchanges doubles — the variant narrows the force evaluation, where the time goes.
#ifdef CBMD_SINGLE_PRECISION
#define CBMD_FORCE_PRECISION 32
#else
#define CBMD_FORCE_PRECISION 64
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
