// neighbor.h — Verlet neighbour list behind the reference's Neighbor surface
// (src/neighbor.h:54-75, src/neighbor_types/neighbor_verlet.h:22-76).
#ifndef CBMD_HOST_NEIGHBOR_H
#define CBMD_HOST_NEIGHBOR_H

#include <cstdint>
#include <vector>

#include "system.h"

// tags standing in for Cabana::FullNeighborTag / HalfNeighborTag and
// Cabana::VerletLayout2D / VerletLayoutCSR
struct FullNeighborTag
{
    static constexpr bool half = false;
};
struct HalfNeighborTag
{
    static constexpr bool half = true;
};
struct VerletLayout2D
{
    static constexpr int layout = CBMD_LAYOUT_2D;
};
struct VerletLayoutCSR
{
    static constexpr int layout = CBMD_LAYOUT_CSR;
};

// Host view of the device list with the three accessors every consumer of a Cabana list
// uses (Cabana::NeighborList<L>::numNeighbor / getNeighbor / maxNeighbor,
// unit_test/tstNeighbor.hpp:64-68).  Filled lazily: the step loop never needs it.
struct VerletListView
{
    cbmd_ctx *ctx = nullptr;
    mutable bool fetched = false;
    mutable std::vector<int> counts, neighbors;
    mutable std::vector<int64_t> offsets;
    mutable int max_n = 0;
    int n_local = 0, n_total = 0;

    void fetch() const
    {
        if ( fetched )
            return;
        int64_t total = 0;
        cbmd_check( cbmd_neigh_sizes( ctx, &total, &max_n ), "cbmd_neigh_sizes" );
        counts.assign( n_total > 0 ? n_total : 1, 0 );
        offsets.assign( n_local + 1, 0 );
        neighbors.assign( total > 0 ? total : 1, 0 );
        cbmd_check( cbmd_neigh_get( ctx, counts.data(), offsets.data(), neighbors.data() ),
                    "cbmd_neigh_get" );
        fetched = true;
    }
    int numNeighbor( int i ) const
    {
        fetch();
        return i < n_total ? counts[i] : 0;
    }
    int getNeighbor( int i, int n ) const
    {
        fetch();
        return neighbors[offsets[i] + n];
    }
    int maxNeighbor() const
    {
        fetch();
        return max_n;
    }
};

template <class t_System>
class Neighbor
{
  public:
    T_X_FLOAT neigh_cut;
    bool half_neigh;
    T_INT max_neigh_guess;

    Neighbor( T_X_FLOAT neigh_cut_, bool half_neigh_, T_INT max_neigh_guess_ = 0 )
        : neigh_cut( neigh_cut_ )
        , half_neigh( half_neigh_ )
        , max_neigh_guess( max_neigh_guess_ )
    {
    }
    virtual ~Neighbor() {}
    virtual void create( t_System *system ) = 0;
    virtual const char *name() { return "Neighbor:None"; }
};

template <class t_System, class t_iteration, class t_layout>
class NeighborVerlet : public Neighbor<t_System>
{
  public:
    using t_neigh_list = VerletListView;

    NeighborVerlet( T_X_FLOAT neigh_cut_, bool half_neigh_, T_INT max_neigh_guess_ = 0 )
        : Neighbor<t_System>( neigh_cut_, half_neigh_, max_neigh_guess_ )
    {
    }

    // rows [0,N_local) over all N_local+N_ghost atoms, d^2 <= neigh_cut^2 inclusive; the
    // list TYPE (full/half) is the compile-time tag, as in the reference where it comes
    // from the command line only (mdfactory.h:79-84)
    void create( t_System *system ) override
    {
        int guess_out = this->max_neigh_guess;
        cbmd_check( cbmd_neigh_build( system->ctx, this->neigh_cut, t_iteration::half ? 1 : 0,
                                      t_layout::layout, this->max_neigh_guess, &guess_out ),
                    "cbmd_neigh_build" );
        this->max_neigh_guess = guess_out; // neighbor_verlet.h:58-61
        list.ctx = system->ctx;
        list.fetched = false;
        list.n_local = system->N_local;
        list.n_total = system->N_local + system->N_ghost;
    }
    t_neigh_list &get() { return list; }
    const char *name() override
    {
        return t_iteration::half ? "Neighbor:CabanaVerletHalf" : "Neighbor:CabanaVerletFull";
    }

  private:
    t_neigh_list list;
};

#endif
