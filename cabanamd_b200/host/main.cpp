// cbnMD — command-line driver (reference bin/main.cpp:57-78): read the options, build
// the application through the factory, init, run.  One process per GPU; rank / size come
// from the launcher's environment (world.h) instead of MPI_Init.
#include <cstdio>
#include <exception>

#include "mdfactory.h"

int main( int argc, char *argv[] )
{
    try
    {
        World::get().init_from_env();
        set_print_rank( World::get().rank == 0 );
        InputCL commandline;
        commandline.read_args( argc, argv );
        if ( !commandline.input_file )
        {
            if ( print_rank() )
                std::fprintf( stderr, "cbnMD: no input deck given (-il FILE)\n" );
            return 1;
        }
        CabanaMD *cabanamd = MDfactory::create( commandline );
        cabanamd->init( commandline );
        cabanamd->run();
        delete cabanamd;
    }
    catch ( const std::exception &e )
    {
        std::fprintf( stderr, "cbnMD (rank %d): %s\n", World::get().rank, e.what() );
        return 1;
    }
    return 0;
}
