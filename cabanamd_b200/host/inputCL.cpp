#include "inputCL.h"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>

#include "output.h"

namespace
{
const char *usage =
    "CabanaMD 0.1 (cabanamd-b200: sm_100a CUDA + NCCL)\nOptions:\n"
    "  -il, --input-lammps [FILE]   LAMMPS-style input deck\n"
    "  -o,  --output-file [FILE]    output file name (default cabanaMD.out)\n"
    "  -e,  --error-file [FILE]     error file name (default cabanaMD.err)\n"
    "  --device-type [TYPE]         CUDA (the only device of this build)\n"
    "  --force-iteration [TYPE]     NEIGH_FULL | NEIGH_HALF\n"
    "  --neigh-parallel [TYPE]      SERIAL | TEAM | TEAM_VECTOR (accepted; one kernel serves all)\n"
    "  --neigh-type [TYPE]          VERLET_2D | VERLET_CSR\n"
    "  --dumpbinary [N] [PATH]      every N steps write PATH/output.<step>.<rank> (binary state)\n"
    "  --correctness [N] [PATH] [FILE]  every N steps compare with PATH/output.* (id-matched\n"
    "                               l2 / max deltas of x, v, f), one line per step in FILE\n"
    "  --vacuum [N]                 enlarge the box N (>1) times\n";

bool is( const char *a, const char *b ) { return std::strcmp( a, b ) == 0; }

// value of the option at argv[i]; a missing value is a usage error
const char *value_of( int argc, char *argv[], int i )
{
    if ( i + 1 >= argc )
        log_err( std::cout, "Missing value for command line option: ", argv[i] );
    return i + 1 < argc ? argv[i + 1] : "";
}

int choose( const std::map<std::string, int> &table, const char *opt, const char *val )
{
    auto it = table.find( val );
    if ( it == table.end() )
        log_err( std::cout, "Unknown commandline option: ", opt, " ", val );
    return it == table.end() ? 0 : it->second;
}
} // namespace

void InputCL::read_args( int argc, char *argv[] )
{
    for ( int i = 1; i < argc; i++ )
    {
        const char *a = argv[i];
        if ( is( a, "-h" ) || is( a, "--help" ) )
            log( std::cout, usage );
        else if ( is( a, "-il" ) || is( a, "--input-lammps" ) )
        {
            input_file = value_of( argc, argv, i++ );
            input_file_type = INPUT_LAMMPS;
        }
        else if ( is( a, "-o" ) || is( a, "--output-file" ) )
            output_file = value_of( argc, argv, i++ );
        else if ( is( a, "-e" ) || is( a, "--error-file" ) )
            error_file = value_of( argc, argv, i++ );
        else if ( is( a, "--device-type" ) )
            device_type = choose( { { "SERIAL", SERIAL }, { "PTHREAD", PTHREAD }, { "OPENMP", OPENMP },
                                    { "CUDA", CUDA }, { "HIP", HIP } },
                                  a, value_of( argc, argv, i++ ) );
        else if ( is( a, "--force-iteration" ) )
        {
            set_force_iteration = true;
            force_iteration_type =
                choose( { { "NEIGH_FULL", FORCE_ITER_NEIGH_FULL }, { "NEIGH_HALF", FORCE_ITER_NEIGH_HALF } },
                        a, value_of( argc, argv, i++ ) );
        }
        else if ( is( a, "--neigh-type" ) )
        {
            neighbor_type = choose( { { "VERLET_2D", NEIGH_VERLET_2D }, { "VERLET_CSR", NEIGH_VERLET_CSR },
                                      { "TREE_2D", NEIGH_TREE_2D }, { "TREE_CSR", NEIGH_TREE_CSR } },
                                    a, value_of( argc, argv, i++ ) );
            if ( neighbor_type == NEIGH_TREE_2D || neighbor_type == NEIGH_TREE_CSR )
                log_err( std::cout, "ArborX requested, but not enabled in Cabana!" );
        }
        else if ( is( a, "--neigh-parallel" ) )
            force_neigh_parallel_type = choose( { { "SERIAL", FORCE_PARALLEL_NEIGH_SERIAL },
                                                  { "TEAM", FORCE_PARALLEL_NEIGH_TEAM },
                                                  { "TEAM_VECTOR", FORCE_PARALLEL_NEIGH_VECTOR } },
                                                a, value_of( argc, argv, i++ ) );
        else if ( is( a, "--dumpbinary" ) )
        {
            dumpbinary_rate = std::atoi( value_of( argc, argv, i ) );
            dumpbinary_path = value_of( argc, argv, i + 1 );
            dumpbinaryflag = true;
            i += 2;
        }
        else if ( is( a, "--correctness" ) )
        {
            correctness_rate = std::atoi( value_of( argc, argv, i ) );
            reference_path = value_of( argc, argv, i + 1 );
            correctness_file = value_of( argc, argv, i + 2 );
            correctnessflag = true;
            i += 3;
        }
        else if ( is( a, "--vacuum" ) )
        {
            vacuum = true;
            vacuum_rate = std::atof( value_of( argc, argv, i++ ) );
            if ( vacuum_rate <= 1.0 )
                log_err( std::cout, "Vacuum multiplier must be bigger than 1.0" );
        }
        else if ( std::strstr( a, "--kokkos-" ) == nullptr )
            log_err( std::cout, "Unknown command line argument: ", a );
    }
}
