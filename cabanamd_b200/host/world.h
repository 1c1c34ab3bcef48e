// world.h — process-level launch information replacing MPI_Init/MPI_COMM_WORLD
// (reference bin/main.cpp:57-78): one process per GPU, rank/size taken from the
// launcher's environment (torchrun-style RANK / WORLD_SIZE / LOCAL_RANK), and the NCCL
// unique id passed through a file named by CBMD_NCCL_ID_FILE (there is no MPI here).
#ifndef CBMD_HOST_WORLD_H
#define CBMD_HOST_WORLD_H

#include <array>
#include <string>

struct World
{
    int rank = 0;
    int nranks = 1;
    int device = 0;
    std::string id_file; // where rank 0 publishes the 128-byte NCCL unique id

    static World &get();
    // read RANK / WORLD_SIZE / LOCAL_RANK / CBMD_NCCL_ID_FILE
    void init_from_env();
    // rank 0: create + publish the id; others: wait for it.  128 bytes.
    void exchange_unique_id( unsigned char id[128] ) const;
    void retire_unique_id() const; // rank 0 removes the published id once all ranks have joined
};

// MPI_Dims_create( n, 3 ): balanced factorisation, non-increasing
// (what Cabana::Grid::DimBlockPartitioner<3>::ranksPerDimension returns, system.h:154-155)
std::array<int, 3> dims_create( int n );

#endif
