// mdfactory.h — run-time options -> compile-time module types (reference
// src/mdfactory.h:58-155), reduced to the one device this build has: the list type
// (full / half) and layout (2D / CSR) come from the command line only, as in the
// reference (mdfactory.h:79-84).
#ifndef CBMD_HOST_MDFACTORY_H
#define CBMD_HOST_MDFACTORY_H

#include <stdexcept>

#include "cabanamd.h"

class MDfactory
{
  public:
    static CabanaMD *create( InputCL commandline )
    {
        const int device = commandline.device_type;
        if ( device != DEFAULT && device != CUDA )
            throw std::runtime_error( "CabanaMD not compiled with requested device type (this build is "
                                      "CUDA sm_100a only, no CPU fallback)" );
        const bool half = commandline.force_iteration_type == FORCE_ITER_NEIGH_HALF;
        switch ( commandline.neighbor_type )
        {
        case NEIGH_VERLET_2D:
            if ( half )
                return new CbnMD<System, NeighborVerlet<System, HalfNeighborTag, VerletLayout2D>>;
            return new CbnMD<System, NeighborVerlet<System, FullNeighborTag, VerletLayout2D>>;
        case NEIGH_VERLET_CSR:
            if ( half )
                return new CbnMD<System, NeighborVerlet<System, HalfNeighborTag, VerletLayoutCSR>>;
            return new CbnMD<System, NeighborVerlet<System, FullNeighborTag, VerletLayoutCSR>>;
        default:
            throw std::runtime_error( "Invalid neighbor type (VERLET_2D and VERLET_CSR are available)" );
        }
    }
};

#endif
