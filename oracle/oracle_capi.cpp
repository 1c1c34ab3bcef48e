// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).
// Plain-C entry points over the CPU restatement so tests/ and bench.py's
// cpu_baseline leg can drive it through ctypes.  Not part of the product.
#include "oracle.hpp"

using namespace orc;

extern "C"
{

// ---- stateless kernels on caller arrays ----------------------------------
// neighbour list: two-call protocol (build -> sizes, then copy out)
struct OrcList
{
    NeighList L;
};

OrcList *orc_list_new() { return new OrcList; }
void orc_list_free( OrcList *l ) { delete l; }
int64_t orc_list_total( OrcList *l ) { return (int64_t)l->L.neigh.size(); }
int orc_list_max( OrcList *l ) { return l->L.max_neigh; }
void orc_list_copy( OrcList *l, int *counts, int64_t *offsets, int *neigh )
{
    std::copy( l->L.counts.begin(), l->L.counts.end(), counts );
    std::copy( l->L.offsets.begin(), l->L.offsets.end(), offsets );
    std::copy( l->L.neigh.begin(), l->L.neigh.end(), neigh );
}
void orc_list_set( OrcList *l, int n_local, int n_total, const int *counts, const int64_t *offsets,
                   const int *neigh )
{
    l->L.counts.assign( counts, counts + n_total );
    l->L.offsets.assign( offsets, offsets + n_local + 1 );
    l->L.neigh.assign( neigh, neigh + offsets[n_local] );
    l->L.max_neigh = 0;
    for ( int i = 0; i < n_local; i++ )
        l->L.max_neigh = std::max( l->L.max_neigh, counts[i] );
}

void orc_neigh_build( OrcList *l, const double *x, int n_local, int n_total, double r, int half,
                      const double *gmin, const double *gmax )
{
    neigh_build( x, n_local, n_total, r, half != 0, gmin, gmax, l->L );
}
void orc_neigh_brute( OrcList *l, const double *x, int n_local, int n_total, double r, int half )
{
    neigh_brute( x, n_local, n_total, r, half != 0, l->L );
}

static Params make_params( int ntypes, const double *mass, const double *lj1, const double *lj2,
                           const double *cutsq )
{
    Params p;
    p.ntypes = ntypes;
    if ( mass )
        p.mass.assign( mass, mass + ntypes );
    else
        p.mass.assign( ntypes, 1.0 );
    if ( lj1 )
    {
        p.lj1.assign( lj1, lj1 + ntypes * ntypes );
        p.lj2.assign( lj2, lj2 + ntypes * ntypes );
        p.cutsq.assign( cutsq, cutsq + ntypes * ntypes );
    }
    return p;
}

// f is accumulated in place (callers zero it first, as the reference does)
void orc_force_lj( OrcList *l, const double *x, const int *type, double *f, int n_local, int half,
                   int ntypes, const double *lj1, const double *lj2, const double *cutsq )
{
    Params p = make_params( ntypes, nullptr, lj1, lj2, cutsq );
    if ( half )
        force_half( x, type, f, n_local, l->L, p );
    else
        force_full( x, type, f, n_local, l->L, p );
}

// full list, float build of the reference (T_X_FLOAT = T_F_FLOAT = float); f accumulated in place
void orc_force_lj_f32( OrcList *l, const double *x, const int *type, double *f, int n_local,
                       int ntypes, const double *lj1, const double *lj2, const double *cutsq )
{
    Params p = make_params( ntypes, nullptr, lj1, lj2, cutsq );
    force_full_f32( x, type, f, n_local, l->L, p );
}

double orc_energy_lj_f32( OrcList *l, const double *x, const int *type, int n_local, int ntypes,
                          const double *lj1, const double *lj2, const double *cutsq )
{
    Params p = make_params( ntypes, nullptr, lj1, lj2, cutsq );
    return energy_full_f32( x, type, n_local, l->L, p );
}

double orc_energy_lj( OrcList *l, const double *x, const int *type, int n_local, int half,
                      int corrected, int ntypes, const double *lj1, const double *lj2,
                      const double *cutsq )
{
    Params p = make_params( ntypes, nullptr, lj1, lj2, cutsq );
    return energy( x, type, n_local, l->L, p, half != 0, corrected != 0 );
}

void orc_integrate( int which, double *x, double *v, const double *f, const int *type, int n_local,
                    int ntypes, const double *mass, double dt, double mvv2e )
{
    Rank r;
    r.N_local = n_local;
    r.x.assign( x, x + 3 * (size_t)n_local );
    r.v.assign( v, v + 3 * (size_t)n_local );
    r.f.assign( f, f + 3 * (size_t)n_local );
    r.type.assign( type, type + n_local );
    Params p = make_params( ntypes, mass, nullptr, nullptr, nullptr );
    p.dt = dt;
    p.mvv2e = mvv2e;
    if ( which == 0 )
        initial_integrate( r, p );
    else
        final_integrate( r, p );
    std::copy( r.x.begin(), r.x.end(), x );
    std::copy( r.v.begin(), r.v.end(), v );
}

// Binning of n_local atoms in a local box [llo,lhi]; returns perm (new[i]=old[perm[i]])
// and the derived grid (nbin[3], bmin[3], bmax[3]).
void orc_binning( const double *x, int n_local, const double *llo, const double *lhi, double dx,
                  double dy, double dz, int halo_depth, int *perm, int *nbin, double *bmin,
                  double *bmax )
{
    Rank r;
    r.resize( n_local );
    r.N_local = n_local;
    std::copy( x, x + 3 * (size_t)n_local, r.x.begin() );
    for ( int d = 0; d < 3; d++ )
    {
        r.dom.llo[d] = llo[d];
        r.dom.lhi[d] = lhi[d];
        r.dom.lext[d] = lhi[d] - llo[d];
    }
    auto pm = create_binning( r, dx, dy, dz, halo_depth );
    std::copy( pm.begin(), pm.end(), perm );
    for ( int d = 0; d < 3; d++ )
    {
        nbin[d] = r.nbin[d];
        bmin[d] = r.bmin[d];
        bmax[d] = r.bmax[d];
    }
}

void orc_velocity_geom( int seed, const double *coord, double *u3 )
{
    RandomVelocityGeom rng;
    rng.reset( seed, coord );
    u3[0] = rng.uniform();
    u3[1] = rng.uniform();
    u3[2] = rng.uniform();
}

// createAtoms of unit_test/tstNeighbor.hpp:262-285 (Kokkos XorShift64 pool, Serial backend)
void orc_kokkos_positions( unsigned long long seed, int n, double lo, double hi, double *x )
{
    kokkos_serial_positions( seed, n, lo, hi, x );
}

void orc_dims_create( int n, int *dims )
{
    auto g = dims_create( n );
    dims[0] = g[0];
    dims[1] = g[1];
    dims[2] = g[2];
}

// ---- the simulation over virtual ranks -----------------------------------
Sim *orc_sim_new( int ntypes, const double *mass, const double *lj1, const double *lj2,
                  const double *cutsq, double force_cutoff, double skin, int half,
                  int exchange_rate, double ghost_cutoff, double dt, double mvv2e, double boltz )
{
    Sim *s = new Sim;
    s->p = make_params( ntypes, mass, lj1, lj2, cutsq );
    s->p.force_cutoff = force_cutoff;
    s->p.skin = skin;
    s->p.half = half != 0;
    s->p.exchange_rate = exchange_rate;
    s->p.ghost_cutoff = ghost_cutoff;
    s->p.dt = dt;
    s->p.mvv2e = mvv2e;
    s->p.boltz = boltz;
    return s;
}
void orc_sim_free( Sim *s ) { delete s; }

void orc_sim_create_lattice_fcc( Sim *s, double lattice_constant, const double *blo,
                                 const double *bhi, int nranks, double temp, int seed )
{
    s->create_lattice_fcc( lattice_constant, blo, bhi, nranks, temp, seed );
}

// Start from caller-provided global atoms (already inside the box): distribute to
// the owning virtual rank by lo <= x < hi (inputFile_impl.h:750-756 ownership rule).
void orc_sim_set_atoms( Sim *s, const double *glo, const double *ghi, int nranks, int n,
                        const double *x, const double *v, const int *type, const int *id )
{
    s->ranks.assign( nranks, Rank() );
    s->N = n;
    for ( int rk = 0; rk < nranks; rk++ )
        s->ranks[rk].dom = make_domain( glo, ghi, nranks, rk, s->p.ghost_cutoff );
    for ( int i = 0; i < n; i++ )
    {
        for ( int rk = 0; rk < nranks; rk++ )
        {
            Rank &r = s->ranks[rk];
            bool in = true;
            for ( int d = 0; d < 3; d++ )
            {
                bool last = r.dom.pos[d] == r.dom.grid[d] - 1;
                in = in && x[3 * i + d] >= r.dom.llo[d] &&
                     ( x[3 * i + d] < r.dom.lhi[d] || ( last && x[3 * i + d] <= r.dom.lhi[d] ) );
            }
            if ( in )
            {
                for ( int d = 0; d < 3; d++ )
                {
                    r.x.push_back( x[3 * i + d] );
                    r.v.push_back( v[3 * i + d] );
                    r.f.push_back( 0.0 );
                }
                r.q.push_back( 0.0 );
                r.type.push_back( type[i] );
                r.id.push_back( id[i] );
                break;
            }
        }
    }
    for ( auto &r : s->ranks )
    {
        r.N_local = r.size();
        r.N_ghost = 0;
    }
}

void orc_sim_setup( Sim *s ) { s->setup(); }
void orc_sim_run( Sim *s, int nsteps, int thermo_rate ) { s->run( nsteps, thermo_rate ); }
int orc_sim_exchange( Sim *s ) { return s->exchange(); }
void orc_sim_binning( Sim *s )
{
    for ( auto &r : s->ranks )
        create_binning( r, s->p.neigh_cut(), s->p.neigh_cut(), s->p.neigh_cut(), 1 );
}
void orc_sim_exchange_halo( Sim *s ) { s->exchange_halo(); }
void orc_sim_update_halo( Sim *s ) { s->update_halo(); }
void orc_sim_neighbor( Sim *s ) { s->neighbor_create(); }
void orc_sim_force( Sim *s ) { s->force_compute(); }
void orc_sim_initial_integrate( Sim *s )
{
    for ( auto &r : s->ranks )
        initial_integrate( r, s->p );
}
void orc_sim_final_integrate( Sim *s )
{
    for ( auto &r : s->ranks )
        final_integrate( r, s->p );
}

int orc_sim_natoms( Sim *s ) { return s->N; }
int orc_sim_nranks( Sim *s ) { return s->nranks(); }
int orc_sim_nlocal( Sim *s, int rk ) { return s->ranks[rk].N_local; }
int orc_sim_nghost( Sim *s, int rk ) { return s->ranks[rk].N_ghost; }
void orc_sim_domain( Sim *s, int rk, double *llo, double *lhi, double *ghost_lo, double *ghost_hi,
                     int *grid, int *pos )
{
    const Domain &d = s->ranks[rk].dom;
    for ( int k = 0; k < 3; k++ )
    {
        llo[k] = d.llo[k];
        lhi[k] = d.lhi[k];
        ghost_lo[k] = d.ghost_lo[k];
        ghost_hi[k] = d.ghost_hi[k];
        grid[k] = d.grid[k];
        pos[k] = d.pos[k];
    }
}
// copies N_local+N_ghost rows
void orc_sim_get( Sim *s, int rk, double *x, double *v, double *f, int *type, int *id )
{
    const Rank &r = s->ranks[rk];
    const size_t n = (size_t)r.N_local + r.N_ghost;
    if ( x )
        std::copy( r.x.begin(), r.x.begin() + 3 * n, x );
    if ( v )
        std::copy( r.v.begin(), r.v.begin() + 3 * n, v );
    if ( f )
        std::copy( r.f.begin(), r.f.begin() + 3 * n, f );
    if ( type )
        std::copy( r.type.begin(), r.type.begin() + n, type );
    if ( id )
        std::copy( r.id.begin(), r.id.begin() + n, id );
}
int64_t orc_sim_list_total( Sim *s, int rk ) { return (int64_t)s->ranks[rk].list.neigh.size(); }
void orc_sim_list_copy( Sim *s, int rk, int *counts, int64_t *offsets, int *neigh )
{
    const NeighList &L = s->ranks[rk].list;
    std::copy( L.counts.begin(), L.counts.end(), counts );
    std::copy( L.offsets.begin(), L.offsets.end(), offsets );
    std::copy( L.neigh.begin(), L.neigh.end(), neigh );
}

double orc_sim_temperature( Sim *s ) { return s->temperature(); }
double orc_sim_kinetic( Sim *s ) { return s->kinetic(); }
double orc_sim_potential( Sim *s, int corrected ) { return s->potential( corrected != 0 ); }
int orc_sim_nthermo( Sim *s ) { return (int)s->thermo.size(); }
void orc_sim_thermo( Sim *s, int k, int *step, double *T, double *PE, double *KE )
{
    *step = s->thermo[k].step;
    *T = s->thermo[k].T;
    *PE = s->thermo[k].PE;
    *KE = s->thermo[k].KE;
}
void orc_sim_record_thermo( Sim *s ) { s->record_thermo(); }
void orc_sim_timers( Sim *s, double *t5 )
{
    t5[0] = s->t_force;
    t5[1] = s->t_neigh;
    t5[2] = s->t_comm;
    t5[3] = s->t_int;
    t5[4] = s->t_other;
}
int orc_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads( int n )
{
#ifdef _OPENMP
    omp_set_num_threads( n );
#else
    (void)n;
#endif
}

} // extern "C"
