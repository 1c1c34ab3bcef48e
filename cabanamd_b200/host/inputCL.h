// inputCL.h — command-line options of cbnMD (reference src/inputCL.h:54-79,
// src/inputCL.cpp:57-258): same flags, same defaults, same error convention.
#ifndef CBMD_HOST_INPUTCL_H
#define CBMD_HOST_INPUTCL_H

#include <string>

#include "types.h"

class InputCL
{
  public:
    int input_file_type = INPUT_LAMMPS;
    int neighbor_type = NEIGH_VERLET_2D;
    int force_iteration_type = FORCE_ITER_NEIGH_FULL;
    bool set_force_iteration = false;
    int force_neigh_parallel_type = FORCE_PARALLEL_NEIGH_SERIAL;
    int device_type = DEFAULT;
    bool vacuum = false;
    double vacuum_rate = 1.0;

    // binary state dumps / regression check (binary_dump.h); InputFile copies them
    int dumpbinary_rate = 0, correctness_rate = 0;
    bool dumpbinaryflag = false, correctnessflag = false;
    const char *dumpbinary_path = nullptr, *reference_path = nullptr, *correctness_file = nullptr;

    const char *input_file = nullptr;
    std::string output_file = "cabanaMD.out";
    std::string error_file = "cabanaMD.err";

    void read_args( int argc, char *argv[] );
};

#endif
