// system.h — particle store + domain scalars behind the reference's System surface
// (src/system.h:65-279, src/system_types/system_1aosoa.h:19-106).
//
// The particle data live on the GPU inside the libcbmd_cuda context (one 32-byte
// position+type record array, SoA v/f); this class owns the context handle, the counts
// N / N_local / N_ghost, the per-type masses, the global / local / ghost mesh scalars and
// host mirrors of the six fields (x v f type id q, row-major [n][3] like the reference's
// slices) that are filled on request.  The reference's three AoSoA layouts are a Kokkos
// tuning knob; they collapse to the one device layout, and name() keeps the log string.
#ifndef CBMD_HOST_SYSTEM_H
#define CBMD_HOST_SYSTEM_H

#include <algorithm>
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cbmd_c_api.h"
#include "types.h"
#include "world.h"

inline void cbmd_check( int rc, const char *what )
{
    if ( rc != 0 )
        throw std::runtime_error( std::string( what ) + ": " + cbmd_last_error() );
}

class System
{
  public:
    T_INT N = 0;       // global particles
    T_INT N_max = 0;   // high-water mark of the storage
    T_INT N_local = 0; // owned
    T_INT N_ghost = 0; // non-owned

    int ntypes = 1;
    std::string atom_style = "atomic";
    std::vector<T_V_FLOAT> mass = std::vector<T_V_FLOAT>( 1, 1.0 ); // per type

    // simulation box, this rank's sub-box and its ghost-mesh bounding box
    T_X_FLOAT global_mesh_x = 0, global_mesh_y = 0, global_mesh_z = 0;
    T_X_FLOAT global_mesh_lo[3] = { 0, 0, 0 }, global_mesh_hi[3] = { 0, 0, 0 };
    T_X_FLOAT local_mesh_x = 0, local_mesh_y = 0, local_mesh_z = 0;
    T_X_FLOAT local_mesh_lo_x = 0, local_mesh_lo_y = 0, local_mesh_lo_z = 0;
    T_X_FLOAT local_mesh_hi_x = 0, local_mesh_hi_y = 0, local_mesh_hi_z = 0;
    T_X_FLOAT ghost_mesh_lo_x = 0, ghost_mesh_lo_y = 0, ghost_mesh_lo_z = 0;
    T_X_FLOAT ghost_mesh_hi_x = 0, ghost_mesh_hi_y = 0, ghost_mesh_hi_z = 0;
    T_X_FLOAT halo_width = 0;
    std::array<int, 3> ranks_per_dim = { 1, 1, 1 };
    std::array<int, 3> rank_dim_pos = { 0, 0, 0 };

    T_FLOAT boltz = 0, mvv2e = 0, dt = 0;

    // host mirrors of the slices (valid after deep_copy_to_host)
    std::vector<T_X_FLOAT> x;
    std::vector<T_V_FLOAT> v;
    std::vector<T_F_FLOAT> f;
    std::vector<T_INT> type, id;
    std::vector<T_FLOAT> q;

    cbmd_ctx *ctx = nullptr;

    System() = default;
    System( const System & ) = delete;
    System &operator=( const System & ) = delete;
    ~System()
    {
        if ( ctx )
            cbmd_destroy( ctx );
    }

    // creates the device context on this rank's GPU (Kokkos::ScopeGuard + `new t_System`)
    void init()
    {
        if ( !ctx )
        {
            cbmd_check( cbmd_create( &ctx, World::get().device ), "cbmd_create" );
            if ( CBMD_FORCE_PRECISION == 32 ) // FP32 build variant (types.h)
                cbmd_check( cbmd_set_option( ctx, "precision", 32 ), "cbmd_set_option(precision)" );
        }
    }

    // host mirrors only grow (system_1aosoa.h:59-67)
    void resize( T_INT N_new )
    {
        if ( N_new > N_max )
            N_max = N_new;
        x.resize( 3 * (size_t)N_new );
        v.resize( 3 * (size_t)N_new );
        f.resize( 3 * (size_t)N_new );
        type.resize( N_new );
        id.resize( N_new );
        q.resize( N_new );
    }

    // the reference re-takes its views after every resize; here the device arrays are
    // owned by the context, so these are no-ops kept for source compatibility
    void slice_all() {}
    void slice_integrate() {}
    void slice_force() {}
    void slice_properties() {}
    void slice_x() {}
    void slice_v() {}
    void slice_f() {}
    void slice_type() {}
    void slice_id() {}
    void slice_q() {}

    // SystemCommon::create_domain (system.h:140-205,251-271): rank grid from
    // MPI_Dims_create, 100 mesh cells per rank per dimension, halo width in cells
    void create_domain( std::array<double, 3> low_corner, std::array<double, 3> high_corner )
    {
        // (sic) the reference mixes indices here, system.h:144-145
        const double ghost_cutoff =
            std::max( std::max( high_corner[0] - low_corner[0], high_corner[2] - low_corner[1] ),
                      high_corner[2] - low_corner[2] );
        create_domain( low_corner, high_corner, ghost_cutoff );
    }
    void create_domain( std::array<double, 3> low_corner, std::array<double, 3> high_corner,
                        double ghost_cutoff )
    {
        const World &w = World::get();
        ranks_per_dim = dims_create( w.nranks );
        // Cartesian coordinates, last dimension fastest (MPI_Cart_create order)
        int r = w.rank;
        rank_dim_pos[2] = r % ranks_per_dim[2];
        r /= ranks_per_dim[2];
        rank_dim_pos[1] = r % ranks_per_dim[1];
        r /= ranks_per_dim[1];
        rank_dim_pos[0] = r;

        const int cells_per_dim_per_rank = 100;
        double cell[3], lo[3], hi[3], glo[3], ghi[3];
        double min_cell = 0;
        for ( int d = 0; d < 3; d++ )
        {
            global_mesh_lo[d] = low_corner[d];
            global_mesh_hi[d] = high_corner[d];
            cell[d] = ( high_corner[d] - low_corner[d] ) / ( cells_per_dim_per_rank * ranks_per_dim[d] );
            min_cell = d == 0 ? cell[d] : std::min( min_cell, cell[d] );
        }
        global_mesh_x = high_corner[0] - low_corner[0];
        global_mesh_y = high_corner[1] - low_corner[1];
        global_mesh_z = high_corner[2] - low_corner[2];
        halo_width = std::ceil( ghost_cutoff / min_cell );
        for ( int d = 0; d < 3; d++ )
        {
            const int off = cells_per_dim_per_rank * rank_dim_pos[d];
            lo[d] = low_corner[d] + cell[d] * off;
            hi[d] = low_corner[d] + cell[d] * ( off + cells_per_dim_per_rank );
            glo[d] = low_corner[d] + cell[d] * ( off - halo_width );
            ghi[d] = low_corner[d] + cell[d] * ( off + cells_per_dim_per_rank + halo_width );
        }
        local_mesh_lo_x = lo[0], local_mesh_lo_y = lo[1], local_mesh_lo_z = lo[2];
        local_mesh_hi_x = hi[0], local_mesh_hi_y = hi[1], local_mesh_hi_z = hi[2];
        ghost_mesh_lo_x = glo[0], ghost_mesh_lo_y = glo[1], ghost_mesh_lo_z = glo[2];
        ghost_mesh_hi_x = ghi[0], ghost_mesh_hi_y = ghi[1], ghost_mesh_hi_z = ghi[2];
        local_mesh_x = hi[0] - lo[0];
        local_mesh_y = hi[1] - lo[1];
        local_mesh_z = hi[2] - lo[2];
        init();
        const int grid[3] = { ranks_per_dim[0], ranks_per_dim[1], ranks_per_dim[2] };
        const int pos[3] = { rank_dim_pos[0], rank_dim_pos[1], rank_dim_pos[2] };
        cbmd_check( cbmd_set_domain( ctx, global_mesh_lo, global_mesh_hi, lo, hi, glo, ghi, grid, pos ),
                    "cbmd_set_domain" );
    }

    // push units + per-type masses to the device tables
    void sync_parameters()
    {
        init();
        cbmd_check( cbmd_set_units( ctx, boltz, mvv2e, dt ), "cbmd_set_units" );
        cbmd_check( cbmd_set_mass( ctx, (int)mass.size(), mass.data() ), "cbmd_set_mass" );
    }

    // host mirrors -> device (System::deep_copy(host_system), inputFile_impl.h:786-791):
    // uploads rows [0,N_local) and drops any ghosts
    void deep_copy_from_host( bool with_forces = false )
    {
        init();
        cbmd_check( cbmd_set_atoms( ctx, N_local, x.data(), v.data(), with_forces ? f.data() : nullptr,
                                    type.data(), id.data(), q.data() ),
                    "cbmd_set_atoms" );
        N_ghost = 0;
    }
    // only the velocities (the rescale after the temperature measurement)
    void deep_copy_velocities_from_host()
    {
        cbmd_check( cbmd_set_velocities( ctx, N_local, v.data() ), "cbmd_set_velocities" );
    }
    // device -> host mirrors, rows [0, N_local (+ N_ghost))
    void deep_copy_to_host( bool with_ghosts = false )
    {
        refresh_counts();
        const T_INT n = N_local + ( with_ghosts ? N_ghost : 0 );
        resize( n );
        if ( n > 0 )
            cbmd_check( cbmd_get_atoms( ctx, 0, n, x.data(), v.data(), f.data(), type.data(), id.data(),
                                        q.data() ),
                        "cbmd_get_atoms" );
    }
    // counts change inside Comm::exchange / exchange_halo
    void refresh_counts()
    {
        cbmd_check( cbmd_get_counts( ctx, &N_local, &N_ghost ), "cbmd_get_counts" );
        if ( N_local + N_ghost > N_max )
            N_max = N_local + N_ghost;
    }

    const char *name() { return "System:1AoSoA"; }
};

#endif
