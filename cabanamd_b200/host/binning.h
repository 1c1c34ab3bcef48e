// binning.h — linked-cell sort behind the reference's Binning surface
// (src/binning_cabana.h:54-68, src/binning_cabana_impl.h:57-113).
#ifndef CBMD_HOST_BINNING_H
#define CBMD_HOST_BINNING_H

#include <stdexcept>

#include "system.h"

template <class t_System>
class Binning
{
    t_System *system;

  public:
    T_INT nbinx = 0, nbiny = 0, nbinz = 0, nhalo = 0;
    T_X_FLOAT minx = 0, maxx = 0, miny = 0, maxy = 0, minz = 0, maxz = 0;

    Binning( t_System *s )
        : system( s )
    {
    }

    // cell-sorts the owned atoms and permutes all six fields.  The reference is only
    // ever called with (do_local, !do_ghost, sort) (cabanamd_impl.h:203-204,312-313);
    // other combinations are rejected rather than silently approximated.
    void create_binning( T_X_FLOAT dx, T_X_FLOAT dy, T_X_FLOAT dz, int halo_depth, bool do_local,
                         bool do_ghost, bool sort )
    {
        if ( !( do_local && !do_ghost && sort ) )
            throw std::runtime_error( "Binning::create_binning: only (do_local, !do_ghost, sort) is "
                                      "supported" );
        int nbin[3];
        double mn[3], mx[3];
        cbmd_check( cbmd_bin_sort( system->ctx, dx, dy, dz, halo_depth, nbin, mn, mx ), "cbmd_bin_sort" );
        nbinx = nbin[0], nbiny = nbin[1], nbinz = nbin[2];
        nhalo = halo_depth;
        minx = mn[0], miny = mn[1], minz = mn[2];
        maxx = mx[0], maxy = mx[1], maxz = mx[2];
    }
    const char *name() { return "Binning:CabanaLinkedCell"; }
};

#endif
