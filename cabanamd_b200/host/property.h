// property.h — thermo properties (src/property_temperature*.h, property_kine*.h,
// property_pote*.h): device sum of m v^2 / pair energy, then a scalar all-reduce.
#ifndef CBMD_HOST_PROPERTY_H
#define CBMD_HOST_PROPERTY_H

#include "comm.h"
#include "force.h"

template <class t_System>
class Temperature
{
    Comm<t_System> *comm;

  public:
    Temperature( Comm<t_System> *c )
        : comm( c )
    {
    }
    // T = mvv2e * sum(m v^2) / ((3N-3) boltz)   (property_temperature_impl.h:55-77)
    T_V_FLOAT compute( t_System *system )
    {
        T_FLOAT T = 0.0;
        cbmd_check( cbmd_sum_mv2( system->ctx, &T ), "cbmd_sum_mv2" );
        const T_INT dof = 3 * system->N - 3;
        const T_V_FLOAT factor = system->mvv2e / ( 1.0 * dof * system->boltz );
        comm->reduce_float( &T, 1 );
        return T * factor;
    }
};

template <class t_System>
class KinE
{
    Comm<t_System> *comm;

  public:
    KinE( Comm<t_System> *c )
        : comm( c )
    {
    }
    // KE = 0.5 mvv2e sum(m v^2)   (property_kine_impl.h:55-76)
    T_V_FLOAT compute( t_System *system )
    {
        T_FLOAT KE = 0.0;
        cbmd_check( cbmd_sum_mv2( system->ctx, &KE ), "cbmd_sum_mv2" );
        const T_V_FLOAT factor = 0.5 * system->mvv2e;
        comm->reduce_float( &KE, 1 );
        return KE * factor;
    }
};

template <class t_System, class t_Neighbor>
class PotE
{
    Comm<t_System> *comm;

  public:
    PotE( Comm<t_System> *c )
        : comm( c )
    {
    }
    // property_pote_impl.h:55-63
    T_F_FLOAT compute( t_System *system, Force<t_System, t_Neighbor> *force, t_Neighbor *neighbor )
    {
        T_F_FLOAT PE = force->compute_energy( system, neighbor );
        comm->reduce_float( &PE, 1 );
        return PE;
    }
};

#endif
