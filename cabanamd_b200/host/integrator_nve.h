// integrator_nve.h — velocity-Verlet NVE behind the reference's Integrator surface
// (src/integrator_nve.h:56-110, src/integrator_nve_impl.h:50-83).
#ifndef CBMD_HOST_INTEGRATOR_NVE_H
#define CBMD_HOST_INTEGRATOR_NVE_H

#include "system.h"

template <class t_System>
class Integrator
{
  public:
    T_V_FLOAT timestep_size;

    Integrator( t_System *s )
        : timestep_size( s->dt )
    {
    }
    // v += dtf/m f; x += dt v   (dtf = 0.5 dt / mvv2e)
    void initial_integrate( t_System *s ) { cbmd_check( cbmd_integrate_initial( s->ctx ), "cbmd_integrate_initial" ); }
    // v += dtf/m f
    void final_integrate( t_System *s ) { cbmd_check( cbmd_integrate_final( s->ctx ), "cbmd_integrate_final" ); }
    const char *name() { return "Integrator:NVE"; }
};

#endif
