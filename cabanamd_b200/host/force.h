// force.h — pair force behind the reference's Force surface (src/force.h:57-74) and the
// LJ implementation (src/force_types/force_lj_cabana_neigh.h, ..._impl.h:62-120,261-377).
#ifndef CBMD_HOST_FORCE_H
#define CBMD_HOST_FORCE_H

#include <cmath>
#include <string>
#include <vector>

#include "system.h"

template <class t_System, class t_Neighbor>
class Force
{
  public:
    Force( t_System * ) {}
    virtual ~Force() {}
    virtual void init_coeff( std::vector<std::vector<std::string>> args ) = 0;
    virtual void compute( t_System *system, t_Neighbor *neighbor ) = 0;
    virtual T_F_FLOAT compute_energy( t_System *, t_Neighbor * ) { return 0.0; }
    virtual const char *name() { return "ForceNone"; }
    // extension: the step loop announces that compute_energy will follow the next
    // compute() at unchanged positions (thermo steps, cabanamd_impl.h:363-366), so the
    // energy can ride along in the same neighbour sweep
    virtual void energy_follows() {}
};

template <class t_System, class t_Neighbor>
class ForceLJ : public Force<t_System, t_Neighbor>
{
    t_System *system;
    int ntypes;
    std::vector<double> lj1, lj2, cutsq;

  public:
    int step = 0;

    ForceLJ( t_System *s )
        : Force<t_System, t_Neighbor>( s )
        , system( s )
        , ntypes( s->ntypes )
        , lj1( (size_t)s->ntypes * s->ntypes, 0.0 )
        , lj2( lj1 )
        , cutsq( lj1 )
    {
    }

    // one `pair_coeff i j eps sigma cut` line per entry (words 1..5), symmetric tables
    // lj1 = 48 eps sigma^12, lj2 = 24 eps sigma^6, cutsq = cut^2 (:62-89)
    void init_coeff( std::vector<std::vector<std::string>> args ) override
    {
        for ( const auto &words : args )
        {
            const int i = std::stoi( words.at( 1 ) ) - 1, j = std::stoi( words.at( 2 ) ) - 1;
            const double eps = std::stod( words.at( 3 ) ), sigma = std::stod( words.at( 4 ) ),
                         cut = std::stod( words.at( 5 ) );
            if ( i < 0 || j < 0 || i >= ntypes || j >= ntypes )
                throw std::runtime_error( "pair_coeff: atom type out of range" );
            lj1[i * ntypes + j] = lj1[j * ntypes + i] = 48.0 * eps * std::pow( sigma, 12.0 );
            lj2[i * ntypes + j] = lj2[j * ntypes + i] = 24.0 * eps * std::pow( sigma, 6.0 );
            cutsq[i * ntypes + j] = cutsq[j * ntypes + i] = cut * cut;
        }
        cbmd_check( cbmd_set_lj( system->ctx, ntypes, lj1.data(), lj2.data(), cutsq.data() ), "cbmd_set_lj" );
    }

    // accumulates into f (the step loop zeroes f first); the force PATH follows
    // neighbor->half_neigh exactly like the reference (:104)
    void compute( t_System *s, t_Neighbor *neighbor ) override
    {
        cbmd_check( cbmd_force_lj( s->ctx, neighbor->half_neigh ? 1 : 0 ), "cbmd_force_lj" );
        step++;
    }
    T_F_FLOAT compute_energy( t_System *s, t_Neighbor *neighbor ) override
    {
        double pe = 0.0;
        cbmd_check( cbmd_energy_lj( s->ctx, neighbor->half_neigh ? 1 : 0, &pe, &last_pe_corrected ),
                    "cbmd_energy_lj" );
        return pe;
    }
    // scalar pair virial of this rank, sum r_ij . f_ij over the pairs inside the cutoff (an extension:
    // the reference's pressure is a TODO, cabanamd_impl.h:477-480); free after energy_follows()
    double compute_virial( t_System *s, t_Neighbor *neighbor )
    {
        double w = 0.0;
        cbmd_check( cbmd_virial_lj( s->ctx, neighbor->half_neigh ? 1 : 0, &w ), "cbmd_virial_lj" );
        return w;
    }
    void energy_follows() override { cbmd_check( cbmd_request_energy( system->ctx ), "cbmd_request_energy" ); }
    const char *name() override { return "Force:LJCabana"; }

    // half-list energy with fac = 1 on every stored pair
    double last_pe_corrected = 0.0;
};

#endif
