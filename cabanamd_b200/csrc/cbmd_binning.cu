// cbmd_binning.cu — counting-sort cell binning and the cell lists used by the
// neighbour build.  Replaces Binning::create_binning (reference
// src/binning_cabana_impl.h:57-113) and the [Cabana] LinkedCellList + permute it
// delegates to (src/system_types/system_1aosoa.h:82-85).
//
// Algorithm: count per cell (atomics; counts are order independent) -> exclusive
// scan -> scatter atom indices (atomic cursors) -> one warp per cell orders its
// slice by ascending index.  The result is the STABLE counting sort, so the
// permutation is deterministic and equal to the oracle's, unlike the reference's
// atomic-arrival order.  The reorder pass then moves all six fields once.
#include <cub/device/device_scan.cuh>

#include "cbmd_internal.cuh"

void cbmd_exclusive_scan_int( cbmd_ctx *ctx, int *data, int n )
{
    // in place over n+1 entries (data[n] must be 0 on entry) so data[n] = total
    size_t tmp = 0;
    CBMD_CUDA( cub::DeviceScan::ExclusiveSum( nullptr, tmp, data, data, n + 1, ctx->stream ) );
    // CUB temp storage belongs to the context (its device, its stream): contexts on other
    // devices or streams of the same thread must not share it; freed by cbmd_destroy
    if ( tmp > ctx->scan_tmp_bytes )
    {
        if ( ctx->scan_tmp )
        {
            CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
            CBMD_CUDA( cudaFree( ctx->scan_tmp ) );
        }
        ctx->scan_tmp = nullptr;
        ctx->scan_tmp_bytes = tmp + ( 1 << 16 );
        CBMD_CUDA( cudaMalloc( &ctx->scan_tmp, ctx->scan_tmp_bytes ) );
    }
    void *d_tmp = ctx->scan_tmp;
    CBMD_CUDA( cub::DeviceScan::ExclusiveSum( d_tmp, tmp, data, data, n + 1, ctx->stream ) );
    ctx->launches += 2;
}

__global__ void __launch_bounds__( 256 )
    k_cell_count( const XT *__restrict__ xt, int first, int n, GridDesc g,
                  int *__restrict__ atom_cell, int *__restrict__ cell_count )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT r = xt[first + i];
    const int c = cell_of( g, r.x, r.y, r.z );
    atom_cell[i] = c;
    atomicAdd( &cell_count[c], 1 );
}

__global__ void __launch_bounds__( 256 )
    k_cell_fill( const int *__restrict__ atom_cell, int first, int n, int *__restrict__ cursor,
                 int *__restrict__ cell_atoms )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const int slot = atomicAdd( &cursor[atom_cell[i]], 1 );
    cell_atoms[slot] = first + i;
}

// Small cells (the half-size cells of the neighbour build hold 2-3 atoms): one LANE per cell,
// insertion sort in place; cells with more than CELL_SORT_SMALL atoms are left to k_cell_sort.
#define CELL_SORT_SMALL 6
__global__ void __launch_bounds__( 256 )
    k_cell_sort_small( const int *__restrict__ cell_start, int ncells, int *__restrict__ cell_atoms )
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if ( c >= ncells )
        return;
    const int b = cell_start[c], n = cell_start[c + 1] - b;
    if ( n <= 1 || n > CELL_SORT_SMALL )
        return;
    int v[CELL_SORT_SMALL];
#pragma unroll
    for ( int k = 0; k < CELL_SORT_SMALL; k++ )
        v[k] = k < n ? cell_atoms[b + k] : 0x7fffffff;
    // sorting network by repeated compare-exchange (odd-even transposition, registers only)
#pragma unroll
    for ( int pass = 0; pass < CELL_SORT_SMALL; pass++ )
#pragma unroll
        for ( int k = pass & 1; k + 1 < CELL_SORT_SMALL; k += 2 )
        {
            const int lo = min( v[k], v[k + 1] ), hi = max( v[k], v[k + 1] );
            v[k] = lo;
            v[k + 1] = hi;
        }
#pragma unroll
    for ( int k = 0; k < CELL_SORT_SMALL; k++ )
        if ( k < n )
            cell_atoms[b + k] = v[k];
}

// one warp per cell: rank each entry by the number of smaller entries (indices are
// unique), then rewrite the slice in ascending order.  small_done: cells of at most
// CELL_SORT_SMALL atoms were already ordered by k_cell_sort_small.
#define CELL_SORT_K 8
__global__ void __launch_bounds__( 256 )
    k_cell_sort( const int *__restrict__ cell_start, int ncells, int *__restrict__ cell_atoms, int small_done )
{
    const int warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int lane = threadIdx.x & 31;
    if ( warp >= ncells )
        return;
    const int b = cell_start[warp], e = cell_start[warp + 1];
    const int n = e - b;
    if ( n <= ( small_done ? CELL_SORT_SMALL : 1 ) )
        return;
    if ( n <= 32 )
    {
        const int mine = lane < n ? cell_atoms[b + lane] : 0x7fffffff;
        int rank = 0;
        for ( int k = 0; k < n; k++ )
        {
            const int other = __shfl_sync( 0xffffffffu, mine, k );
            rank += other < mine;
        }
        if ( lane < n )
            cell_atoms[b + rank] = mine;
        return;
    }
    if ( n <= 32 * CELL_SORT_K )
    {
        int val[CELL_SORT_K], rank[CELL_SORT_K];
#pragma unroll
        for ( int s = 0; s < CELL_SORT_K; s++ )
        {
            const int k = lane + 32 * s;
            val[s] = k < n ? cell_atoms[b + k] : 0x7fffffff;
            rank[s] = 0;
        }
        for ( int k = 0; k < n; k++ )
        {
            const int other = cell_atoms[b + k];
#pragma unroll
            for ( int s = 0; s < CELL_SORT_K; s++ )
                rank[s] += other < val[s];
        }
        __syncwarp();
#pragma unroll
        for ( int s = 0; s < CELL_SORT_K; s++ )
            if ( lane + 32 * s < n )
                cell_atoms[b + rank[s]] = val[s];
        return;
    }
    // very crowded cell (not expected for a liquid): serial insertion sort
    if ( lane == 0 )
        for ( int a = b + 1; a < e; a++ )
        {
            const int key = cell_atoms[a];
            int p = a - 1;
            while ( p >= b && cell_atoms[p] > key )
            {
                cell_atoms[p + 1] = cell_atoms[p];
                p--;
            }
            cell_atoms[p + 1] = key;
        }
}

void cbmd_build_cell_lists_grid( cbmd_ctx *ctx, const GridDesc &g, int first, int count )
{
    const int ncells = g.n[0] * g.n[1] * g.n[2];
    if ( ncells + 1 > ctx->ncells_cap )
    {
        if ( ctx->cell_start )
        {
            CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
            CBMD_CUDA( cudaFree( ctx->cell_start ) );
            CBMD_CUDA( cudaFree( ctx->cell_cursor ) );
        }
        ctx->ncells_cap = ncells + ncells / 4 + 64;
        CBMD_CUDA( cudaMalloc( &ctx->cell_start, (size_t)ctx->ncells_cap * sizeof( int ) ) );
        CBMD_CUDA( cudaMalloc( &ctx->cell_cursor, (size_t)ctx->ncells_cap * sizeof( int ) ) );
    }
    cudaStream_t s = ctx->stream;
    CBMD_CUDA( cudaMemsetAsync( ctx->cell_start, 0, (size_t)( ncells + 1 ) * sizeof( int ), s ) );
    if ( count > 0 )
    {
        k_cell_count<<<div_up( count, 256 ), 256, 0, s>>>( ctx->xt, first, count, g,
                                                           ctx->atom_cell, ctx->cell_start );
        CBMD_LAUNCH_CHECK( ctx );
    }
    cbmd_exclusive_scan_int( ctx, ctx->cell_start, ncells );
    CBMD_CUDA( cudaMemcpyAsync( ctx->cell_cursor, ctx->cell_start, (size_t)ncells * sizeof( int ),
                                cudaMemcpyDeviceToDevice, s ) );
    if ( count > 0 )
    {
        k_cell_fill<<<div_up( count, 256 ), 256, 0, s>>>( ctx->atom_cell, first, count,
                                                          ctx->cell_cursor, ctx->cell_atoms );
        CBMD_LAUNCH_CHECK( ctx );
        // mean occupancy decides: grids of small cells get the lane-per-cell pass first
        const int small = (double)count < 8.0 * (double)ncells ? 1 : 0;
        if ( small )
        {
            k_cell_sort_small<<<div_up( ncells, 256 ), 256, 0, s>>>( ctx->cell_start, ncells, ctx->cell_atoms );
            CBMD_LAUNCH_CHECK( ctx );
        }
        k_cell_sort<<<div_up( ncells, 8 ), 256, 0, s>>>( ctx->cell_start, ncells, ctx->cell_atoms, small );
        CBMD_LAUNCH_CHECK( ctx );
    }
}

// gather-permute of all six fields: new[i] = old[perm[i]]
__global__ void __launch_bounds__( 256 )
    k_permute( const int *__restrict__ perm, int n, int cap, const XT *__restrict__ xt,
               XT *__restrict__ xt_o, const double *__restrict__ v, double *__restrict__ v_o,
               const double *__restrict__ f, double *__restrict__ f_o, const int *__restrict__ id,
               int *__restrict__ id_o, const double *__restrict__ q, double *__restrict__ q_o )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const int o = perm[i];
    xt_o[i] = ld_xt( xt + o );
#pragma unroll
    for ( int c = 0; c < 3; c++ )
    {
        v_o[(size_t)c * cap + i] = v[(size_t)c * cap + o];
        f_o[(size_t)c * cap + i] = f[(size_t)c * cap + o];
    }
    id_o[i] = id[o];
    q_o[i] = q[o];
}

// derive the grid of create_binning (binning_cabana_impl.h:74-99 + [Cabana]
// LinkedCellList: n = floor((max-min)/delta), dx' = (max-min)/n)
void cbmd_binning_grid( const cbmd_ctx *ctx, const double din[3], int halo_depth, int nbin[3],
                        double bmin[3], double bmax[3], GridDesc &g )
{
    double delta[3];
    for ( int d = 0; d < 3; d++ )
    {
        const double ext = ctx->lhi[d] - ctx->llo[d];
        nbin[d] = (int)( ext / din[d] );
        if ( nbin[d] == 0 )
            nbin[d] = 1;
        delta[d] = ext / nbin[d];
    }
    const double eps = delta[0] / 1000;
    for ( int d = 0; d < 3; d++ )
    {
        bmin[d] = -delta[d] * halo_depth - eps + ctx->llo[d];
        bmax[d] = delta[d] * halo_depth + eps + ctx->lhi[d];
        int n = (int)floor( ( bmax[d] - bmin[d] ) / delta[d] );
        if ( n < 1 )
            n = 1;
        const double dx = ( bmax[d] - bmin[d] ) / n;
        g.mn[d] = bmin[d];
        g.rdx[d] = 1.0 / dx;
        g.n[d] = n;
    }
}

extern "C" int cbmd_bin_sort( cbmd_ctx *ctx, double dx, double dy, double dz, int halo_depth,
                              int nbin_out[3], double min_out[3], double max_out[3] )
{
    CBMD_API_BEGIN
    TimedRegion timed__( ctx, CBMD_T_OTHER );
    CBMD_REQUIRE( ctx->have_domain, "cbmd_set_domain must be called before cbmd_bin_sort" );
    CBMD_REQUIRE( dx > 0 && dy > 0 && dz > 0, "bin sizes must be positive" );
    cbmd_materialize_zero_force( ctx );
    cbmd_bump_epoch( ctx, true, true );
    const double din[3] = { dx, dy, dz };
    GridDesc g;
    cbmd_binning_grid( ctx, din, halo_depth, ctx->nbin, ctx->bmin, ctx->bmax, g );
    ctx->nhalo = halo_depth;
    for ( int d = 0; d < 3; d++ )
    {
        ctx->ncell[d] = g.n[d];
        ctx->brdx[d] = g.rdx[d];
        if ( nbin_out )
            nbin_out[d] = ctx->nbin[d];
        if ( min_out )
            min_out[d] = ctx->bmin[d];
        if ( max_out )
            max_out[d] = ctx->bmax[d];
    }
    ctx->have_bins = true;
    const int n = ctx->n_local;
    // binning acts on the owned atoms only (do_local && !do_ghost)
    cbmd_build_cell_lists_grid( ctx, g, 0, n );
    if ( n > 0 )
    {
        cudaStream_t s = ctx->stream;
        CBMD_CUDA( cudaMemcpyAsync( ctx->perm, ctx->cell_atoms, (size_t)n * sizeof( int ),
                                    cudaMemcpyDeviceToDevice, s ) );
        k_permute<<<div_up( n, 256 ), 256, 0, s>>>( ctx->perm, n, ctx->cap, ctx->xt, ctx->xt_alt,
                                                    ctx->v, ctx->v_alt, ctx->f, ctx->f_alt,
                                                    ctx->id, ctx->id_alt, ctx->q, ctx->q_alt );
        CBMD_LAUNCH_CHECK( ctx );
        // ghosts (if any) sit after the owned block and are not permuted: carry them over
        const int ng = ctx->n_ghost;
        if ( ng > 0 )
        {
            CBMD_CUDA( cudaMemcpyAsync( ctx->xt_alt + n, ctx->xt + n, (size_t)ng * sizeof( XT ),
                                        cudaMemcpyDeviceToDevice, s ) );
            CBMD_CUDA( cudaMemcpyAsync( ctx->id_alt + n, ctx->id + n, (size_t)ng * sizeof( int ),
                                        cudaMemcpyDeviceToDevice, s ) );
            CBMD_CUDA( cudaMemcpyAsync( ctx->q_alt + n, ctx->q + n, (size_t)ng * sizeof( double ),
                                        cudaMemcpyDeviceToDevice, s ) );
            for ( int c = 0; c < 3; c++ )
            {
                CBMD_CUDA( cudaMemcpyAsync( ctx->v_alt + (size_t)c * ctx->cap + n,
                                            ctx->v + (size_t)c * ctx->cap + n,
                                            (size_t)ng * sizeof( double ),
                                            cudaMemcpyDeviceToDevice, s ) );
                CBMD_CUDA( cudaMemcpyAsync( ctx->f_alt + (size_t)c * ctx->cap + n,
                                            ctx->f + (size_t)c * ctx->cap + n,
                                            (size_t)ng * sizeof( double ),
                                            cudaMemcpyDeviceToDevice, s ) );
            }
        }
        std::swap( ctx->xt, ctx->xt_alt );
        std::swap( ctx->v, ctx->v_alt );
        std::swap( ctx->f, ctx->f_alt );
        std::swap( ctx->id, ctx->id_alt );
        std::swap( ctx->q, ctx->q_alt );
    }
    ctx->perm_n = n;
    // a sort invalidates index-based derived state
    ctx->nb_n = 0;
    ctx->nb_ntot = 0;
    if ( ctx->n_ghost > 0 )
        ctx->have_halo = false;
    CBMD_API_END
}

extern "C" int cbmd_get_permutation( cbmd_ctx *ctx, int *perm )
{
    CBMD_API_BEGIN
    CBMD_REQUIRE( perm != nullptr, "null output" );
    if ( ctx->perm_n > 0 )
    {
        CBMD_CUDA( cudaMemcpyAsync( perm, ctx->perm, (size_t)ctx->perm_n * sizeof( int ),
                                    cudaMemcpyDeviceToHost, ctx->stream ) );
        CBMD_CUDA( cudaStreamSynchronize( ctx->stream ) );
    }
    CBMD_API_END
}
