// cbmd_neighbor.cu — Verlet neighbour-list build.
// Replaces NeighborVerlet::create (reference src/neighbor_types/neighbor_verlet.h:43-62)
// and the [Cabana] VerletList<…, Full|Half, 2D|CSR, TeamVectorOpTag> it constructs.
//
// Semantics kept bit-exactly: row i in [0,n_local) holds every j in
// [0,n_local+n_ghost) with isValid(i,j) and dx*dx+dy*dy+dz*dz <= r*r (inclusive,
// evaluated un-contracted, left to right); Full: i != j; Half: i != j and
// (xj>xi || (xj==xi && (yj>yi || (yj==yi && zj>zi)))); ghost rows are empty.
//
// Device layout: padded 2-D table, addressed by nb_entry() (cbmd_internal.cuh): tiles of 32
// atoms, four entries of a row per 16 bytes, so a warp of the one-lane-per-atom sweeps reads
// four neighbours per lane as one coalesced 512-byte request and a tile's rows are one
// contiguous block.  Rows are padded to a multiple of four with the atom's own index.  The
// CSR view the reference also offers is produced on demand by cbmd_neigh_get.  Row capacity
// follows Cabana's 2-D policy: start from max_neigh_guess (rounded up to a multiple of four),
// and if any row overflows rebuild at 1.1 x the observed maximum.
#include "cbmd_internal.cuh"

void cbmd_build_cell_lists_grid( cbmd_ctx *ctx, const GridDesc &g, int first, int count );
void cbmd_binning_grid( const cbmd_ctx *ctx, const double din[3], int halo_depth, int nbin[3],
                        double bmin[3], double bmax[3], GridDesc &g );

__device__ __forceinline__ bool half_valid( const XT &a, const XT &b )
{
    return b.x > a.x || ( b.x == a.x && ( b.y > a.y || ( b.y == a.y && b.z > a.z ) ) );
}

// ---------------------------------------------------------------------------
// k_neigh_build: one CTA per NBC z-consecutive cells of one (x,y) column of the Verlet
// grid, one warp per cell, one LANE per owned atom of that cell.
//
//  1. The 27-cell stencils of the CTA's cells are together at most nine runs that are
//     contiguous in cell_atoms (z is the fastest cell index).  The CTA stages the
//     candidates of those runs once in shared memory as 16-byte records
//     {x,y,z relative to the column block centre as FP32, atom index}.
//  2. Each warp walks the candidates of its own cell's stencil (nine sub-segments of
//     the staged runs); the record is a shared-memory BROADCAST read, and every lane
//     tests it against its own atom, so no cross-lane compaction is needed: a lane
//     appends straight to its row of the transposed table.
//  3. d^2 <= r^2 is decided in FP32 only when that is provably equal to the exact
//     answer: with tol0 bounding the relative FP32 error (> 10x the worst case) the
//     candidate is "surely inside" below r2lo = (r2-tol0)/(1+tol0) and "surely outside"
//     above r2hi = (r2+tol0)/(1-tol0); anything in between is re-evaluated with the
//     reference's exact FP64 expression (un-contracted, left to right) from the global
//     positions, so the SET is bit-identical to the oracle's.  The half-list
//     discriminator (xj > xi, ties by y, z) gets the same treatment.
// Stencils larger than the staging buffer are processed in chunks; per-lane counts
// live in registers across chunks.  Rows are in ascending (cell, index) order.
// ---------------------------------------------------------------------------
// K staged candidates against this lane's atom: FP32 decisions, ONE warp vote for the
// (rare) exact FP64 re-evaluation, then in-order, branch-free appends to the lane's row.
// MODE 0: full list (i != j).  MODE 1: half list (reference HalfNeighborTag).  MODE 2: PULL rows
// for the atomics-free Newton-3 sweep (cbmd_force.cu): the row of an OWNED atom i holds its half
// row (x_j > x_i, any j) plus, flagged with NB_JSIDE, the owned atoms j < i in that order whose
// half row holds i; the row of a GHOST atom holds only the flagged kind.  Stripped of the
// flagged entries the rows ARE the reference's half list (cbmd_neigh_get does that).
template <int MODE, int K>
__device__ __forceinline__ void
sweep_group( const float4 *__restrict__ cand, const XT *__restrict__ xt, const XT &xi, float xr,
             float yr, float zr, int i, float r2lo, float r2hi, float tolx, double rsqr,
             char *row0, int nb_rows, int &count, int n_local, int &count_i )
{
    // i < 0 marks an inactive lane: its r2lo/r2hi are -1 so nothing is ever accepted
    float4 c[K];
#pragma unroll
    for ( int k = 0; k < K; k++ )
        c[k] = cand[k];
    bool ok[K], amb[K];
    int val[K]; // the entry to store: index (+ NB_JSIDE)
    bool any_amb = false;
    const bool iown = i < n_local;
    // MODE 2, per lane: an owned atom keeps x_j > x_i (i side) and owned x_j < x_i (j side), a
    // ghost atom only the latter.  Folded into thresholds so that a candidate costs two compares:
    // "up" can never hold for a ghost lane.
    const float inf = __int_as_float( 0x7f800000 );
    const float up_sure = iown ? tolx : inf, up_maybe = iown ? -tolx : inf;
#pragma unroll
    for ( int k = 0; k < K; k++ )
    {
        const float fx = c[k].x - xr, fy = c[k].y - yr, fz = c[k].z - zr;
        const float d2 = fx * fx + fy * fy + fz * fz;
        const int j = __float_as_int( c[k].w );
        const bool ns = j != i;
        ok[k] = ns & ( d2 < r2lo );
        amb[k] = ns & ( d2 < r2hi );
        val[k] = j;
        if ( MODE == 1 )
        {
            // xj > xi decided in FP32 unless |xj - xi| is within its error
            ok[k] = ok[k] && ( fx > tolx );
            amb[k] = amb[k] && ( fx >= -tolx );
        }
        if ( MODE == 2 )
        {
            // (bitwise on purpose: no short-circuit branches in the candidate loop)
            const bool jown = j < n_local;
            const bool dn = jown & ( fx < -tolx ); // surely the j side
            ok[k] = ok[k] & ( ( fx > up_sure ) | dn );
            amb[k] = amb[k] & ( ( fx >= up_maybe ) | ( jown & ( fx <= tolx ) ) );
            val[k] = dn ? ( j | NB_JSIDE ) : j;
        }
        amb[k] = amb[k] & !ok[k];
        any_amb = any_amb | amb[k];
    }
    if ( __any_sync( 0xffffffffu, any_amb ) )
    {
#pragma unroll
        for ( int k = 0; k < K; k++ )
            if ( amb[k] )
            { // exact re-evaluation of the reference criterion (rare)
                const int j = __float_as_int( c[k].w );
                const XT xj = ld_xt( xt + j );
                ok[k] = dist2_exact( __dsub_rn( xi.x, xj.x ), __dsub_rn( xi.y, xj.y ),
                                     __dsub_rn( xi.z, xj.z ) ) <= rsqr;
                if ( MODE == 1 )
                    ok[k] = ok[k] && half_valid( xi, xj );
                if ( MODE == 2 )
                {
                    const bool up = half_valid( xi, xj ), dn = half_valid( xj, xi );
                    ok[k] = ok[k] && ( ( iown && up ) || ( j < n_local && dn ) );
                    val[k] = dn ? ( j | NB_JSIDE ) : j;
                }
            }
    }
#pragma unroll
    for ( int k = 0; k < K; k++ )
    {
        // predicated store (no branch): entries 4k..4k+3 of this lane's row are one int4,
        // consecutive int4s of the row are 512 bytes apart
        const int st = ( ok[k] && count < nb_rows ) ? 1 : 0;
        const unsigned off = ( ( (unsigned)count >> 2 ) << 9 ) + ( ( (unsigned)count & 3u ) << 2 );
        char *dst = row0 + (unsigned long long)off;
        asm volatile( "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.b32 [%0], %1;\n\t}"
                      :
                      : "l"( dst ), "r"( val[k] ), "r"( st )
                      : "memory" );
        count += ok[k] ? 1 : 0;
        if ( MODE == 2 )
            count_i += ( ok[k] && !( val[k] & NB_JSIDE ) ) ? 1 : 0;
    }
}

#define NBC 4
#define NB_THREADS ( 32 * NBC )
#define NB_STAGE 1280 // candidates per staging chunk (20 KB)

template <int MODE>
__global__ void __launch_bounds__( NB_THREADS )
    k_neigh_build( const XT *__restrict__ xt, int n_local, GridDesc g,
                   const int *__restrict__ cell_start, const int *__restrict__ cell_atoms,
                   double rsqr, double3 centre, int *__restrict__ nb, int nb_stride, int nb_rows,
                   int *__restrict__ nb_count, int *__restrict__ d_max, int *__restrict__ nb_count_i )
{
    __shared__ float4 cand[NB_STAGE];
    __shared__ int run_src[9], run_off[10];
    __shared__ int s_mag;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nz = g.n[2], nzb = ( nz + NBC - 1 ) / NBC;
    const int col = blockIdx.x / nzb, c0 = ( blockIdx.x - col * nzb ) * NBC;
    const int ca = col / g.n[1], cb = col - ca * g.n[1];
    const int c1 = min( c0 + NBC, nz ) - 1; // last cell of this block
    const int colrow = ( ca * g.n[1] + cb ) * nz;

    // this warp's cell and its owned atoms (owned atoms come first inside a cell:
    // ascending index, ghosts have index >= n_local)
    const int cc = c0 + warp;
    int cs = 0, ce = 0;
    if ( cc <= c1 )
    {
        cs = cell_start[colrow + cc];
        ce = cell_start[colrow + cc + 1];
    }
    // rows are made for the owned atoms; the pull rows also for the ghosts
    const bool has_owned = ( ce > cs ) && ( MODE == 2 || cell_atoms[cs] < n_local );
    if ( !__syncthreads_or( has_owned ) )
        return;

    // nine runs (x,y neighbours of the column) x cells [c0-1, c1+1]
    const int zlo = max( c0 - 1, 0 ), zhi = min( c1 + 1, nz - 1 );
    int seg_b = 0, seg_e = 0; // lane r < 9: this warp's sub-segment of run r (staged coords)
    {
        int s0 = 0, len = 0, w0 = 0, w1 = 0;
        if ( lane < 9 )
        {
            const int a = ca + lane / 3 - 1, b = cb + lane % 3 - 1;
            if ( a >= 0 && a < g.n[0] && b >= 0 && b < g.n[1] )
            {
                const int row = ( a * g.n[1] + b ) * nz;
                s0 = cell_start[row + zlo];
                len = cell_start[row + zhi + 1] - s0;
                if ( cc <= c1 )
                {
                    w0 = cell_start[row + max( cc - 1, 0 )] - s0;
                    w1 = cell_start[row + min( cc + 1, nz - 1 ) + 1] - s0;
                }
            }
        }
        int off = len; // inclusive scan over lanes 0..8
        for ( int o = 1; o < 16; o <<= 1 )
        {
            const int v = __shfl_up_sync( 0xffffffffu, off, o );
            if ( lane >= o )
                off += v;
        }
        off -= len; // exclusive
        seg_b = off + w0;
        seg_e = off + w1;
        if ( warp == 0 )
        {
            if ( lane < 9 )
            {
                run_src[lane] = s0;
                run_off[lane] = off;
            }
            if ( lane == 9 )
                run_off[9] = off; // == total (len is 0 for lanes >= 9)
            if ( lane == 0 )
                s_mag = 0;
        }
    }
    __syncthreads();
    const int total = run_off[9];

    // origin of the FP32 relative coordinates: centre of this block of cells
    const double ox = g.rdx[0] > 0.0 ? g.mn[0] + ( ca + 0.5 ) / g.rdx[0] : centre.x;
    const double oy = g.rdx[1] > 0.0 ? g.mn[1] + ( cb + 0.5 ) / g.rdx[1] : centre.y;
    const double oz = g.rdx[2] > 0.0 ? g.mn[2] + ( 0.5 * ( c0 + c1 ) + 0.5 ) / g.rdx[2] : centre.z;
    const float r2f = (float)rsqr;

    // passes of 32 atoms of each warp's cell (one pass unless a cell is crowded); the
    // pass and chunk loops are CTA-uniform because staging needs block barriers
    __shared__ int s_npass;
    if ( threadIdx.x == 0 )
        s_npass = 1;
    __syncthreads();
    if ( lane == 0 && ce - cs > 32 )
        atomicMax( &s_npass, ( ce - cs + 31 ) >> 5 );
    __syncthreads();
    const int n_pass = s_npass;
    const bool multi_chunk = total > NB_STAGE;

    for ( int pass = 0; pass < n_pass; pass++ )
    {
        const int s = cs + pass * 32 + lane;
        int i = -1;
        if ( s < ce )
        {
            i = cell_atoms[s];
            if ( MODE != 2 && i >= n_local )
                i = -1;
        }
        const bool active = i >= 0;
        const bool any_active = __any_sync( 0xffffffffu, active );
        XT xi;
        xi.x = xi.y = xi.z = 0.0;
        xi.t = 0;
        if ( active )
            xi = ld_xt( xt + i );
        const float xr = active ? (float)( xi.x - ox ) : 0.f, yr = active ? (float)( xi.y - oy ) : 0.f,
                    zr = active ? (float)( xi.z - oz ) : 0.f;
        char *const row0 = (char *)( nb + nb_tile_base( active ? i : 0, nb_rows ) );
        int count = 0, count_i = 0;

        for ( int chunk = 0; chunk < total; chunk += NB_STAGE )
        {
            const int nstage = min( NB_STAGE, total - chunk );
            // ---- stage [chunk, chunk + nstage) of the concatenated runs; a single chunk
            //      is staged once and stays valid for later passes
            if ( multi_chunk || pass == 0 )
            {
                if ( multi_chunk && ( chunk > 0 || pass > 0 ) )
                    __syncthreads(); // everyone is done with the previous contents
                float mag = 0.f;
                for ( int t = threadIdx.x; t < nstage; t += NB_THREADS )
                {
                    const int q = chunk + t;
                    int r = 0;
#pragma unroll
                    for ( int k = 1; k < 9; k++ )
                        r += ( q >= run_off[k] );
                    const int j = cell_atoms[run_src[r] + ( q - run_off[r] )];
                    const XT xj = ld_xt( xt + j );
                    const float4 c = make_float4( (float)( xj.x - ox ), (float)( xj.y - oy ),
                                                  (float)( xj.z - oz ), __int_as_float( j ) );
                    mag = fmaxf( mag, fmaxf( fabsf( c.x ), fmaxf( fabsf( c.y ), fabsf( c.z ) ) ) );
                    cand[t] = c;
                }
                // largest |relative coordinate| staged so far (non-negative floats order like ints)
                for ( int o = 16; o > 0; o >>= 1 )
                    mag = fmaxf( mag, __shfl_xor_sync( 0xffffffffu, mag, o ) );
                if ( lane == 0 )
                    atomicMax( &s_mag, __float_as_int( mag ) );
                __syncthreads();
            }
            if ( !any_active )
                continue;
            // FP32 error of d2: relative coordinates carry <= 2^-24*M each (M = largest
            // magnitude), so |fx - dx| <= 2.4e-7*M and, with three more roundings in the
            // sum, |d2f - d2| <= (8.4e-7*M + 1.8e-7)*(1 + d2).  tol0 is > 10x that.
            float M = fmaxf( fabsf( xr ), fmaxf( fabsf( yr ), fabsf( zr ) ) );
            for ( int o = 16; o > 0; o >>= 1 )
                M = fmaxf( M, __shfl_xor_sync( 0xffffffffu, M, o ) );
            M = fmaxf( M, __int_as_float( s_mag ) );
            const float tol0 = 4.0e-6f + 1.0e-5f * M;
            const float r2lo = ( r2f - tol0 ) / ( 1.0f + tol0 ) * 0.999999f;
            const float r2hi = ( r2f + tol0 ) / ( 1.0f - tol0 ) * 1.000001f;
            const float tolx = 1.0e-5f * fmaxf( 1.0f, M );
            // inactive lanes accept nothing
            const float r2lo_l = active ? r2lo : -1.0f, r2hi_l = active ? r2hi : -1.0f;
#pragma unroll 1
            for ( int r = 0; r < 9; r++ )
            {
                const int b = max( __shfl_sync( 0xffffffffu, seg_b, r ) - chunk, 0 );
                const int e = max( min( __shfl_sync( 0xffffffffu, seg_e, r ) - chunk, nstage ), b );
                const int n4 = ( e - b ) >> 2;
                const float4 *cp = cand + b;
                for ( int q = 0; q < n4; q++, cp += 4 )
                    sweep_group<MODE, 4>( cp, xt, xi, xr, yr, zr, i, r2lo_l, r2hi_l, tolx, rsqr,
                                          row0, nb_rows, count, n_local, count_i );
                for ( int t = b + 4 * n4; t < e; t++ )
                    sweep_group<MODE, 1>( cand + t, xt, xi, xr, yr, zr, i, r2lo_l, r2hi_l, tolx,
                                          rsqr, row0, nb_rows, count, n_local, count_i );
            }
        }
        if ( active )
        {
            nb_count[i] = count;
            if ( MODE == 2 )
                nb_count_i[i] = count_i;
            // pad the row to a multiple of four with the atom itself (never a neighbour)
            for ( int k = count; k < min( ( count + 3 ) & ~3, nb_rows ); k++ )
                *(int *)( row0 + ( ( (unsigned)k >> 2 ) << 9 ) + ( ( (unsigned)k & 3u ) << 2 ) ) = i;
        }
        int mx = count, mi = count_i;
        for ( int o = 16; o > 0; o >>= 1 )
        {
            mx = max( mx, __shfl_xor_sync( 0xffffffffu, mx, o ) );
            mi = max( mi, __shfl_xor_sync( 0xffffffffu, mi, o ) );
        }
        if ( lane == 0 && mx > 0 )
        {
            atomicMax( d_max, mx );
            if ( MODE == 2 )
                atomicMax( d_max + 1, mi ); // longest reference row (without the j-side entries)
        }
    }
}

// ---------------------------------------------------------------------------
// k_neigh_build_walk (option "neigh_kernel" 2, the default, and 1): one THREAD per atom walks its
// own cell stencil.
//
// k_neigh_build (option 0, round 1) gives a warp one cell and shares a staged 27-cell stencil
// between the cell's atoms: ~19 of 32 lanes have an atom, and every lane pays the address
// arithmetic of an append for every candidate — 2.74 G warp instructions per build at 4 M atoms,
// 85 % issue-bound (profiles/r2_neigh_kernels.txt).  Here a warp is 32 consecutive atoms (all
// lanes busy; in the MD loop they are neighbours in space because the atoms are cell-sorted),
// each thread streams the runs of ITS stencil — the columns of the cell grid are contiguous in
// cell order, and the candidates are packed once per build in that order as 16-byte records
// {x,y,z relative to the box centre as FP32, atom index} (k_pack_candidates), so a run is a
// sequence of LDG.128 that the lanes of a warp share in L1 — and works in two phases per 32
// candidates: DECIDE (FP32 test, one bit per candidate, no stores; fully unrolled) and APPEND
// (only the accepted ones, ~1 in 7).  The decision is k_neigh_build's: FP32 with an explicit
// error bound (M = the largest relative coordinate, measured while packing), exact FP64
// re-evaluation of the reference criterion from the 32-byte records for the few ambiguous
// candidates — the SET is bit-identical to the oracle's.
//
// reach 1 (option 2): cells >= r, 3x3x3 stencil, 9 runs of ~58 candidates.  Rows ascend in atom
// index like k_neigh_build's (the atoms are sorted by these very cells), which is the order the
// force sweep is fastest on.  Measured on B200, 4 M atoms: build 3.2 -> 2.7 ms, force sweep
// unchanged, 3.82e9 -> 3.93e9 atom-steps/s.
// reach 2 (option 1): cells >= r/2, 5x5x5 stencil, 25 runs of ~12: 40 % fewer candidates and the
// fastest build (2.0 ms + 0.3 ms for the finer cell lists), but its rows come out in half-cell
// order and the force sweep is 4 % slower on them (0.717 against 0.687 ms) — 20 sweeps lose more
// than one build gains, so it is only the A/B leg.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_pack_candidates( const XT *__restrict__ xt, const int *__restrict__ cell_atoms, int n_total,
                       double3 origin, float4 *__restrict__ cpos, int *__restrict__ mag_bits )
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    float mag = 0.f;
    if ( s < n_total )
    {
        const int j = cell_atoms[s];
        const XT r = ld_xt( xt + j );
        const float4 c = make_float4( (float)( r.x - origin.x ), (float)( r.y - origin.y ),
                                      (float)( r.z - origin.z ), __int_as_float( j ) );
        cpos[s] = c;
        mag = fmaxf( fabsf( c.x ), fmaxf( fabsf( c.y ), fabsf( c.z ) ) );
    }
    for ( int o = 16; o > 0; o >>= 1 )
        mag = fmaxf( mag, __shfl_xor_sync( 0xffffffffu, mag, o ) );
    if ( ( threadIdx.x & 31 ) == 0 && mag > 0.f )
        atomicMax( mag_bits, __float_as_int( mag ) ); // non-negative floats order like ints
}

// MODE as in sweep_group: 0 full, 1 half, 2 PULL rows (rows for ghost atoms too, n_rows = all atoms)
template <int MODE>
__global__ void __launch_bounds__( 128 )
    k_neigh_build_walk( const XT *__restrict__ xt, int n_local, int n_rows, GridDesc g,
                        const int *__restrict__ cell_start, const float4 *__restrict__ cpos,
                        const int *__restrict__ atom_cell, double3 origin, double rsqr,
                        const int *__restrict__ mag_bits, int *__restrict__ nb, int nb_rows,
                        int *__restrict__ nb_count, int *__restrict__ d_max, int reach,
                        int *__restrict__ nb_count_i )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int count = 0, count_i = 0;
    if ( i < n_rows )
    {
        // FP32 error of d2: relative coordinates carry <= 2^-24*M each, so |fx - dx| <= 1.2e-7*M
        // and, with three more roundings in the sum, |d2f - d2| <= (8.4e-7*M + 1.8e-7)*(1 + d2);
        // tol0 is > 10x that (same bound as k_neigh_build)
        const float M = __int_as_float( *mag_bits );
        const float r2f = (float)rsqr;
        const float tol0 = 4.0e-6f + 1.0e-5f * M;
        const float r2lo = ( r2f - tol0 ) / ( 1.0f + tol0 ) * 0.999999f;
        const float r2hi = ( r2f + tol0 ) / ( 1.0f - tol0 ) * 1.000001f;
        const float tolx = 1.0e-5f * fmaxf( 1.0f, M );
        const XT xi = ld_xt( xt + i );
        const float xr = (float)( xi.x - origin.x ), yr = (float)( xi.y - origin.y ), zr = (float)( xi.z - origin.z );
        // PULL rows: an owned atom keeps x_j > x_i (i side) and owned x_j < x_i (j side, flagged),
        // a ghost atom only the latter; folded into thresholds ("up" never holds for a ghost)
        const bool iown = i < n_local;
        const float inf = __int_as_float( 0x7f800000 );
        const float up_sure = iown ? tolx : inf, up_maybe = iown ? -tolx : inf;
        const int c = atom_cell[i];
        const int nz = g.n[2], ny = g.n[1], nx = g.n[0];
        const int cz = c % nz, cy = ( c / nz ) % ny, cx = c / ( nz * ny );
        const int zlo = max( cz - reach, 0 ), zhi = min( cz + reach, nz - 1 );
        const int alo = max( cx - reach, 0 ), ahi = min( cx + reach, nx - 1 );
        const int blo = max( cy - reach, 0 ), bhi = min( cy + reach, ny - 1 );
        char *const row0 = (char *)( nb + nb_tile_base( i, nb_rows ) );
        for ( int a = alo; a <= ahi; a++ )
            for ( int b = blo; b <= bhi; b++ )
            {
                const int row = ( a * ny + b ) * nz;
                const int s0 = __ldg( cell_start + row + zlo ), s1 = __ldg( cell_start + row + zhi + 1 );
                // the run in chunks of 32 candidates: first DECIDE (one bit per candidate, no
                // stores), then APPEND the accepted ones — the address arithmetic of an append is
                // not paid for the rejected candidates.  sure / amb: surely accepted / to be
                // decided exactly; js: accepted entries that are the j side of their pair (PULL).
                for ( int c0 = s0; c0 < s1; c0 += 32 )
                {
                    const int rem = min( 32, s1 - c0 );
                    const float4 *cp = cpos + c0;
                    unsigned sure = 0u, amb = 0u, js = 0u;
#define CBMD_DECIDE( K )                                                                          \
    {                                                                                             \
        const float4 q = __ldg( cp + ( K ) );                                                     \
        const float fx = q.x - xr, fy = q.y - yr, fz = q.z - zr;                                  \
        const float d2 = fx * fx + fy * fy + fz * fz;                                             \
        bool in_lo = d2 < r2lo, in_hi = d2 < r2hi;                                                \
        if ( MODE == 1 )                                                                          \
        { /* xj > xi decided in FP32 unless |xj - xi| is within its error */                      \
            in_lo = in_lo & ( fx > tolx );                                                        \
            in_hi = in_hi & ( fx >= -tolx );                                                      \
        }                                                                                         \
        if ( MODE == 2 )                                                                          \
        {                                                                                         \
            const bool jown = __float_as_int( q.w ) < n_local;                                    \
            const bool dn = jown & ( fx < -tolx );                                                \
            in_lo = in_lo & ( ( fx > up_sure ) | dn );                                            \
            in_hi = in_hi & ( ( fx >= up_maybe ) | ( jown & ( fx <= tolx ) ) );                   \
            js |= dn ? ( 1u << ( K ) ) : 0u;                                                      \
        }                                                                                         \
        sure |= in_lo ? ( 1u << ( K ) ) : 0u;                                                     \
        amb |= in_hi ? ( 1u << ( K ) ) : 0u;                                                      \
    }
                    if ( rem >= 12 )
                    {
                        // fully unrolled (the bits are immediates).  A short chunk is decided as a
                        // whole one and the bits past its end are dropped: what lies there are the
                        // next cells' candidates (cpos has slack behind the last atom)
#pragma unroll
                        for ( int k = 0; k < 32; k++ )
                            CBMD_DECIDE( k )
                        const unsigned live = rem == 32 ? 0xffffffffu : ( ( 1u << rem ) - 1u );
                        sure &= live;
                        amb &= live;
                    }
                    else
                    {
#pragma unroll 1
                        for ( int k = 0; k < rem; k++ )
                            CBMD_DECIDE( k )
                    }
#undef CBMD_DECIDE
                    amb &= ~sure;
                    while ( amb )
                    { // exact re-evaluation of the reference criterion (rare)
                        const int k = __ffs( amb ) - 1;
                        amb &= amb - 1u;
                        const int j = __float_as_int( __ldg( cp + k ).w );
                        const XT xj = ld_xt( xt + j );
                        bool ok = dist2_exact( __dsub_rn( xi.x, xj.x ), __dsub_rn( xi.y, xj.y ), __dsub_rn( xi.z, xj.z ) ) <= rsqr;
                        if ( MODE == 1 )
                            ok = ok && half_valid( xi, xj );
                        if ( MODE == 2 )
                        {
                            const bool up = half_valid( xi, xj ), dn = half_valid( xj, xi );
                            ok = ok && ( ( iown && up ) || ( j < n_local && dn ) );
                            js = dn ? ( js | ( 1u << k ) ) : ( js & ~( 1u << k ) );
                        }
                        sure |= ok ? ( 1u << k ) : 0u;
                    }
                    while ( sure )
                    {
                        const int k = __ffs( sure ) - 1;
                        sure &= sure - 1u;
                        const int j = __float_as_int( __ldg( &cp[k].w ) );
                        if ( j != i ) // the atom itself passes the distance test
                        {
                            const bool jside = MODE == 2 && ( ( js >> k ) & 1u );
                            if ( count < nb_rows )
                                *(int *)( row0 + ( ( (unsigned)count >> 2 ) << 9 ) + ( ( (unsigned)count & 3u ) << 2 ) ) =
                                    jside ? ( j | NB_JSIDE ) : j;
                            count++;
                            count_i += jside ? 0 : 1;
                        }
                    }
                }
            }
        nb_count[i] = count;
        if ( MODE == 2 )
            nb_count_i[i] = count_i;
        // pad the row to a multiple of four with the atom itself (never a neighbour)
        for ( int k = count; k < min( ( count + 3 ) & ~3, nb_rows ); k++ )
            *(int *)( row0 + ( ( (unsigned)k >> 2 ) << 9 ) + ( ( (unsigned)k & 3u ) << 2 ) ) = i;
    }
    int mx = count, mi = count_i;
    for ( int o = 16; o > 0; o >>= 1 )
    {
        mx = max( mx, __shfl_xor_sync( 0xffffffffu, mx, o ) );
        mi = max( mi, __shfl_xor_sync( 0xffffffffu, mi, o ) );
    }
    if ( ( threadIdx.x & 31 ) == 0 && mx > 0 )
    {
        atomicMax( d_max, mx );
        if ( MODE == 2 )
            atomicMax( d_max + 1, mi ); // longest reference row (without the j-side entries)
    }
}

// ---------------------------------------------------------------------------
// k_rows_bank_order (option "row_order" 1; NOT the default, see the end of this comment):
// bank-aware order of every row.
//
// The full-list sweeps gather x_j with LDG.128: L1 serves such a request in groups of eight
// lanes, and a group costs as many passes as its worst conflict on the POSITION of the 16-byte
// word inside its 128-byte line, j & 7 (round 1, experiments/ldg_patterns.cu).  Rows in index
// order leave that to chance.  Here lane q (= i & 7) of a group takes, at step r, a neighbour
// of class (r + q) & 7: a Latin square — the eight lanes of a group read eight different
// positions for as long as all classes last; when a class runs out the entry comes from the
// class with most entries left.  The SET of a row does not change, only its order (sums
// differ by round-off from the index-ordered row's; every run orders the same way).
// Measured on B200, 4 M atoms (profiles/r2_force_variants_b4.txt, profiles/r2_row_order.txt):
// FP64 sweep 0.688 -> 0.658 ms (roofline 0.322 -> 0.337), FP32 sweep 0.469 -> 0.407 ms (0.439 ->
// 0.506).  The pass itself takes 2.4 ms per rebuild (three dependent per-lane walks over the
// staged row), 20 sweeps gain 0.6 ms (FP64) / 1.2 ms (FP32): a net loss at the reference's
// rebuild period of 20 steps, so the default keeps the rows in index order.
//
// One warp per 32-atom tile, whole tile block staged in shared memory: bucket every lane's
// row by class (stable), deal it out in Latin order, write the block back coalesced.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__( 128 )
    k_rows_bank_order( int4 *__restrict__ nb4, const int *__restrict__ nb_count, int rows4, int n_local )
{
    extern __shared__ int sm[]; // [4 warps][2][rows4*4][32]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * 4 + w;
    if ( tile * 32 >= n_local )
        return;
    const int i = tile * 32 + lane;
    const int rows = rows4 * 4;
    int *s0 = sm + (size_t)w * 2 * rows * 32, *s1 = s0 + rows * 32;
    const int c = i < n_local ? min( nb_count[i], rows ) : 0;
    int cmax = c;
    for ( int o = 16; o > 0; o >>= 1 )
        cmax = max( cmax, __shfl_xor_sync( 0xffffffffu, cmax, o ) );
    const int c4max = ( cmax + 3 ) >> 2;
    int4 *p = nb4 + ( (size_t)tile * rows4 ) * 32 + lane;
    const int q = i & 7;
    // 1. load the block (coalesced), column of lane l at s0[n*32 + l]; class sizes in packed bytes
    //    (class relative to the lane: (j - q) & 7; classes 0-3 in lo, 4-7 in hi)
    unsigned lo = 0u, hi = 0u;
    for ( int k4 = 0; k4 < c4max; k4++ )
    {
        const int4 v = p[k4 * 32];
        const int jj[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
        {
            const int n = 4 * k4 + u;
            s0[n * 32 + lane] = jj[u];
            if ( n < c )
            {
                const int cl = ( jj[u] - q ) & 7;
                if ( cl < 4 )
                    lo += 1u << ( 8 * cl );
                else
                    hi += 1u << ( 8 * ( cl - 4 ) );
            }
        }
    }
    // 2. stable bucket by class into s1: bucket start = sum of the sizes of the lower classes
    unsigned off_lo, off_hi; // running write offsets per class, packed bytes
    {
        unsigned acc = 0u, t;
        off_lo = 0u;
#pragma unroll
        for ( int b = 0; b < 4; b++ )
        {
            off_lo |= acc << ( 8 * b );
            t = ( lo >> ( 8 * b ) ) & 0xffu;
            acc += t;
        }
        off_hi = 0u;
#pragma unroll
        for ( int b = 0; b < 4; b++ )
        {
            off_hi |= acc << ( 8 * b );
            t = ( hi >> ( 8 * b ) ) & 0xffu;
            acc += t;
        }
    }
    const unsigned start_lo = off_lo, start_hi = off_hi;
    for ( int n = 0; n < c; n++ )
    {
        const int j = s0[n * 32 + lane];
        const int cl = ( j - q ) & 7;
        const unsigned sh = 8 * ( cl & 3 );
        const unsigned o = ( ( cl < 4 ? off_lo : off_hi ) >> sh ) & 0xffu;
        s1[o * 32 + lane] = j;
        if ( cl < 4 )
            off_lo += 1u << sh;
        else
            off_hi += 1u << sh;
    }
    // 3. deal out: step r wants relative class r & 7; exhausted -> the class with most left
    unsigned used_lo = 0u, used_hi = 0u; // entries taken per class
    for ( int r = 0; r < c; r++ )
    {
        int cl = r & 7;
        unsigned sh = 8 * ( cl & 3 );
        unsigned left = ( ( ( cl < 4 ? lo : hi ) >> sh ) & 0xffu ) - ( ( ( cl < 4 ? used_lo : used_hi ) >> sh ) & 0xffu );
        if ( left == 0u )
        {
            // per-class remaining counts, byte-wise; pick the largest (lowest class on ties)
            const unsigned rl = __vsub4( lo, used_lo ), rh = __vsub4( hi, used_hi );
            unsigned best = 0u;
            cl = 0;
#pragma unroll
            for ( int b = 0; b < 8; b++ )
            {
                const unsigned v = ( ( b < 4 ? rl : rh ) >> ( 8 * ( b & 3 ) ) ) & 0xffu;
                if ( v > best )
                {
                    best = v;
                    cl = b;
                }
            }
            sh = 8 * ( cl & 3 );
        }
        const unsigned st = ( ( cl < 4 ? start_lo : start_hi ) >> sh ) & 0xffu;
        const unsigned us = ( ( cl < 4 ? used_lo : used_hi ) >> sh ) & 0xffu;
        s0[r * 32 + lane] = s1[( st + us ) * 32 + lane];
        if ( cl < 4 )
            used_lo += 1u << sh;
        else
            used_hi += 1u << sh;
    }
    // (the padding entries c .. 4*ceil(c/4) of s0 still hold the atom's own index)
    __syncwarp();
    const int c4 = ( c + 3 ) >> 2;
    for ( int k4 = 0; k4 < c4; k4++ )
        p[k4 * 32] = make_int4( s0[( 4 * k4 ) * 32 + lane], s0[( 4 * k4 + 1 ) * 32 + lane],
                                s0[( 4 * k4 + 2 ) * 32 + lane], s0[( 4 * k4 + 3 ) * 32 + lane] );
}

__global__ void __launch_bounds__( 256 )
    k_nb_to_csr( const int *__restrict__ nb, int nb_rows,
                 const int *__restrict__ nb_count, const int64_t *__restrict__ offsets, int n_local,
                 int *__restrict__ csr )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n_local )
        return;
    const int c = nb_count[i];
    int64_t o = offsets[i];
    for ( int n = 0; n < c; n++ )
    {
        const int e = nb[nb_entry( i, n, nb_rows )];
        if ( !( e & NB_JSIDE ) ) // pull rows: the flagged entries are not part of the reference row
            csr[o++] = e;
    }
}

// ---------------------------------------------------------------------------
// Interior / boundary tiles for the halo-compute overlap.  A ghost lies outside (or on)
// the faces of the owned box, so an owned atom farther than r from every face has no
// ghost in its row; a tile (32 consecutive atoms) whose atoms are all that far inside
// can be processed while the ghost positions are still in flight.  The margin is
// widened by 1e-9 relative so rounding can only make the test more conservative.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__( 256 )
    k_tile_flags( const XT *__restrict__ xt, int n_local, double3 lo, double3 hi, double r,
                  int *__restrict__ flag, int n_tiles )
{
    const int tile = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int lane = threadIdx.x & 31;
    if ( tile >= n_tiles )
        return;
    const int i = tile * 32 + lane;
    bool near = false;
    if ( i < n_local )
    {
        const XT a = ld_xt( xt + i );
        near = ( a.x - lo.x <= r ) || ( hi.x - a.x <= r ) || ( a.y - lo.y <= r ) || ( hi.y - a.y <= r ) ||
               ( a.z - lo.z <= r ) || ( hi.z - a.z <= r );
    }
    const unsigned any = __ballot_sync( 0xffffffffu, near );
    if ( lane == 0 )
        flag[tile] = any ? 1 : 0;
    if ( tile == 0 && lane == 0 )
        flag[n_tiles] = 0;
}

// pos = exclusive scan of flag: interior tiles first (ascending), then boundary tiles
__global__ void __launch_bounds__( 256 )
    k_tile_partition( const int *__restrict__ flag, const int *__restrict__ pos, int n_tiles,
                      int *__restrict__ list )
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if ( t >= n_tiles )
        return;
    const int nb = pos[n_tiles]; // number of boundary tiles
    const int p = pos[t];
    if ( flag[t] )
        list[( n_tiles - nb ) + p] = t;
    else
        list[t - p] = t;
}

static void build_tile_lists( cbmd_ctx *ctx, double rcut )
{
    ctx->tiles_valid = false;
    const int n_tiles = ( ctx->n_local + 31 ) >> 5;
    if ( n_tiles == 0 || ctx->flat_halo_ok || !ctx->overlap || !ctx->have_halo )
        return;
    cudaStream_t s = ctx->stream;
    if ( n_tiles + 1 > ctx->tile_cap )
    {
        if ( ctx->tile_list )
        {
            CBMD_CUDA( cudaStreamSynchronize( s ) );
            CBMD_CUDA( cudaFree( ctx->tile_list ) );
            CBMD_CUDA( cudaFree( ctx->tile_flag ) );
        }
        ctx->tile_cap = n_tiles + n_tiles / 4 + 64;
        CBMD_CUDA( cudaMalloc( &ctx->tile_list, (size_t)ctx->tile_cap * sizeof( int ) ) );
        CBMD_CUDA( cudaMalloc( &ctx->tile_flag, (size_t)ctx->tile_cap * sizeof( int ) ) );
    }
    int *pos = (int *)cbmd_scratch( ctx, (size_t)( n_tiles + 1 ) * sizeof( int ) );
    const double r = rcut * ( 1.0 + 1e-9 );
    k_tile_flags<<<div_up( n_tiles * 32, 256 ), 256, 0, s>>>(
        ctx->xt, ctx->n_local, make_double3( ctx->llo[0], ctx->llo[1], ctx->llo[2] ),
        make_double3( ctx->lhi[0], ctx->lhi[1], ctx->lhi[2] ), r, ctx->tile_flag, n_tiles );
    CBMD_LAUNCH_CHECK( ctx );
    CBMD_CUDA( cudaMemcpyAsync( pos, ctx->tile_flag, (size_t)( n_tiles + 1 ) * sizeof( int ),
                                cudaMemcpyDeviceToDevice, s ) );
    cbmd_exclusive_scan_int( ctx, pos, n_tiles );
    k_tile_partition<<<div_up( n_tiles, 256 ), 256, 0, s>>>( ctx->tile_flag, pos, n_tiles,
                                                             ctx->tile_list );
    CBMD_LAUNCH_CHECK( ctx );
    CBMD_CUDA( cudaMemcpyAsync( ctx->h_pinned_i + 2, pos + n_tiles, sizeof( int ),
                                cudaMemcpyDeviceToHost, s ) );
    CBMD_CUDA( cudaStreamSynchronize( s ) );
    ctx->n_tiles_boundary = ctx->h_pinned_i[2];
    ctx->n_tiles_interior = n_tiles - ctx->n_tiles_boundary;
    ctx->tiles_n_local = ctx->n_local;
    ctx->tiles_rcut = rcut;
    ctx->tiles_valid = true;
}

extern "C" int cbmd_neigh_build( cbmd_ctx *ctx, double rcut, int half, int layout,
                                 int max_neigh_guess, int *max_neigh_guess_out )
{
    CBMD_API_BEGIN
    TimedRegion timed__( ctx, CBMD_T_NEIGH );
    CBMD_REQUIRE( ctx->have_domain, "cbmd_set_domain must be called before cbmd_neigh_build" );
    CBMD_REQUIRE( rcut > 0, "neighbour cutoff must be positive" );
    CBMD_REQUIRE( layout == CBMD_LAYOUT_2D || layout == CBMD_LAYOUT_CSR, "unknown list layout" );
    const int n_local = ctx->n_local, n_total = ctx->n_local + ctx->n_ghost;
    cudaStream_t s = ctx->stream;
    cbmd_bump_epoch( ctx, false, false );

    // cell grid of size >= rcut around the owned box; atoms outside are clamped into
    // the edge cells, which keeps |cell(i)-cell(j)| <= 1 for every pair within rcut.
    // option "neigh_kernel" 2 (default): cells >= rcut, one halo layer, 3x3x3 stencil walked by one
    // thread per atom; 1: cells >= rcut/2, two halo layers, 5x5x5 stencil, same kernel; 0: warp
    // per cell over a staged 27-cell stencil (round 1)
    // half lists for the atomics-free Newton-3 sweep get PULL rows (sweep_group MODE 2): rows for
    // owned and ghost atoms, built by the staged kernel
    const bool pull = half && ctx->half_kernel == 1;
    const bool walk = ctx->neigh_kernel >= 1;
    const bool walk_half = ctx->neigh_kernel == 1; // 1: half-size cells, 5x5x5; 2: cells >= rcut, 3x3x3
    const double cell = ( walk && walk_half ) ? 0.5 * rcut : rcut;
    const double din[3] = { cell, cell, cell };
    int nbin[3];
    double gmin[3], gmax[3];
    GridDesc g;
    cbmd_binning_grid( ctx, din, ( walk && walk_half ) ? 2 : 1, nbin, gmin, gmax, g );
    for ( int d = 0; d < 3; d++ )
        if ( 1.0 / g.rdx[d] < din[d] )
        {
            // owned box thinner than the cell: one cell spans this dimension
            g.n[d] = 1;
            g.rdx[d] = 0.0;
        }
    cbmd_build_cell_lists_grid( ctx, g, 0, n_total );

    ctx->nb_half = half ? 1 : 0;
    ctx->nb_layout = layout;
    ctx->nb_rcut = rcut;
    ctx->nb_n = n_local;
    ctx->nb_ntot = n_total;
    ctx->nb_pull = pull;
    int rows = ( ( max_neigh_guess > 0 ? max_neigh_guess : 1 ) + 3 ) & ~3;
    if ( pull ) // a pull row holds both sides of the atom's pairs: about twice the half row
        rows = std::max( ctx->pull_rows_hint, ( 2 * rows + 8 + 3 ) & ~3 );
    const int stride = ( ( pull ? n_total : n_local ) + 31 ) & ~31;
    const double rsqr = rcut * rcut;
    const double3 centre = make_double3( 0.5 * ( ctx->llo[0] + ctx->lhi[0] ),
                                         0.5 * ( ctx->llo[1] + ctx->lhi[1] ),
                                         0.5 * ( ctx->llo[2] + ctx->lhi[2] ) );
    int *d_max = ctx->d_flags; // [0] longest row, [1] longest reference row of a pull table
    int *d_mag = ctx->d_flags + 9;
    if ( n_total > 0 )
    {
        CBMD_CUDA( cudaMemsetAsync( ctx->nb_count, 0, (size_t)n_total * sizeof( int ), s ) );
        if ( pull )
            CBMD_CUDA( cudaMemsetAsync( ctx->nb_count_i, 0, (size_t)n_total * sizeof( int ), s ) );
    }
    if ( walk && n_total > 0 )
    {
        // candidates packed in cell order: {x,y,z relative to the box centre as FP32, index}
        if ( ctx->cpos_cap < ctx->cap )
        {
            if ( ctx->cpos )
            {
                CBMD_CUDA( cudaStreamSynchronize( s ) );
                CBMD_CUDA( cudaFree( ctx->cpos ) );
            }
            ctx->cpos = nullptr;
            // + 32 records of slack: the decide phase may read (and ignore) a whole chunk past the end
            CBMD_CUDA( cudaMalloc( &ctx->cpos, ( (size_t)ctx->cap + 32 ) * sizeof( float4 ) ) );
            CBMD_CUDA( cudaMemsetAsync( ctx->cpos, 0, ( (size_t)ctx->cap + 32 ) * sizeof( float4 ), s ) );
            ctx->cpos_cap = ctx->cap;
        }
        CBMD_CUDA( cudaMemsetAsync( d_mag, 0, sizeof( int ), s ) );
        k_pack_candidates<<<div_up( n_total, 256 ), 256, 0, s>>>( ctx->xt, ctx->cell_atoms, n_total, centre,
                                                                  ctx->cpos, d_mag );
        CBMD_LAUNCH_CHECK( ctx );
    }
    int observed = 0;
    for ( int attempt = 0; attempt < 3; attempt++ )
    {
        const size_t need = nb_table_size( stride > 0 ? stride : 32, rows );
        if ( need > ctx->nb_alloc )
        {
            if ( ctx->nb )
            {
                CBMD_CUDA( cudaStreamSynchronize( s ) );
                if ( ctx->tex_nb )
                    CBMD_CUDA( cudaDestroyTextureObject( ctx->tex_nb ) );
                ctx->tex_nb = 0;
                CBMD_CUDA( cudaFree( ctx->nb ) );
            }
            ctx->nb = nullptr;
            ctx->nb_alloc = ( need + need / 8 + 3 ) & ~(size_t)3;
            CBMD_CUDA( cudaMalloc( &ctx->nb, ctx->nb_alloc * sizeof( int ) ) );
            // the table as 16-byte texels: index stream of the FP32 sweep (a linear texture
            // addresses at most 2^27 texels; larger tables are read through LDG.128)
            if ( ctx->nb_alloc / 4 <= ( (size_t)1 << 27 ) )
            {
                cudaResourceDesc rd = {};
                rd.resType = cudaResourceTypeLinear;
                rd.res.linear.devPtr = ctx->nb;
                rd.res.linear.desc = cudaCreateChannelDesc<int4>();
                rd.res.linear.sizeInBytes = ctx->nb_alloc * sizeof( int );
                cudaTextureDesc td = {};
                td.readMode = cudaReadModeElementType;
                CBMD_CUDA( cudaCreateTextureObject( &ctx->tex_nb, &rd, &td, nullptr ) );
            }
        }
        ctx->nb_rows = rows;
        ctx->nb_stride = stride;
        CBMD_CUDA( cudaMemsetAsync( d_max, 0, 2 * sizeof( int ), s ) );
        if ( n_local > 0 && walk )
        {
            const int n_rows = pull ? n_total : n_local, reach = walk_half ? 2 : 1;
#define NBW_LAUNCH( MODE )                                                                        \
    k_neigh_build_walk<MODE><<<div_up( n_rows, 128 ), 128, 0, s>>>(                               \
        ctx->xt, n_local, n_rows, g, ctx->cell_start, ctx->cpos, ctx->atom_cell, centre, rsqr, d_mag, ctx->nb, rows, \
        ctx->nb_count, d_max, reach, ctx->nb_count_i )
            if ( pull )
                NBW_LAUNCH( 2 );
            else if ( half )
                NBW_LAUNCH( 1 );
            else
                NBW_LAUNCH( 0 );
#undef NBW_LAUNCH
            CBMD_LAUNCH_CHECK( ctx );
        }
        else if ( n_local > 0 )
        {
            const int blocks = g.n[0] * g.n[1] * ( ( g.n[2] + NBC - 1 ) / NBC );
#define NB_LAUNCH( MODE )                                                                         \
    k_neigh_build<MODE><<<blocks, NB_THREADS, 0, s>>>( ctx->xt, n_local, g, ctx->cell_start,      \
                                                       ctx->cell_atoms, rsqr, centre, ctx->nb,    \
                                                       stride, rows, ctx->nb_count, d_max,        \
                                                       ctx->nb_count_i )
            if ( pull )
                NB_LAUNCH( 2 );
            else if ( half )
                NB_LAUNCH( 1 );
            else
                NB_LAUNCH( 0 );
#undef NB_LAUNCH
            CBMD_LAUNCH_CHECK( ctx );
        }
        // NeighborList<>::maxNeighbor (neighbor_verlet.h:58-59)
        CBMD_CUDA( cudaMemcpyAsync( ctx->h_pinned_i, d_max, 2 * sizeof( int ), cudaMemcpyDeviceToHost,
                                    s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
        observed = ctx->h_pinned_i[0];
        if ( observed <= rows )
        {
            if ( pull )
            {
                ctx->pull_rows_hint = std::max( ctx->pull_rows_hint, ( observed + observed / 8 + 3 ) & ~3 );
                observed = ctx->h_pinned_i[1]; // what the reference's maxNeighbor reports
            }
            break;
        }
        rows = (int)( observed * 1.1 ); // [Cabana] 2-D regrow + refill
        if ( rows < observed )
            rows = observed;
        rows = ( rows + 3 ) & ~3;
        CBMD_REQUIRE( attempt < 2, "neighbour list did not converge" );
    }
    ctx->nb_max = observed;
    // bank-aware row order for the LDG.128 gathers of the full-list sweeps (byte counters: rows
    // of at most 255 entries; the shared-memory staging bounds the row capacity as well)
    if ( ctx->row_order == 1 && !half && n_local > 0 && rows <= 252 )
    {
        const size_t smb = (size_t)4 * 2 * rows * 32 * sizeof( int );
        if ( smb <= 200 * 1024 )
        {
            if ( smb > 48 * 1024 )
                CBMD_CUDA( cudaFuncSetAttribute( k_rows_bank_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb ) );
            k_rows_bank_order<<<div_up( div_up( n_local, 32 ), 4 ), 128, smb, s>>>( (int4 *)ctx->nb, ctx->nb_count, rows >> 2,
                                                                                n_local );
            CBMD_LAUNCH_CHECK( ctx );
        }
    }
    build_tile_lists( ctx, rcut );
    // neighbor_verlet.h:60-61: max_neigh_guess = current_max * 1.1 when exceeded
    int guess = max_neigh_guess;
    if ( observed > guess )
        guess = (int)( observed * 1.1 );
    if ( max_neigh_guess_out )
        *max_neigh_guess_out = guess;
    CBMD_API_END
}

extern "C" int cbmd_neigh_get( cbmd_ctx *ctx, int *counts, int64_t *offsets, int *neighbors )
{
    CBMD_API_BEGIN
    const int n_local = ctx->nb_n, n_total = ctx->nb_ntot;
    CBMD_REQUIRE( n_total == ctx->n_local + ctx->n_ghost && n_local == ctx->n_local,
                  "neighbour list is stale (atoms changed since cbmd_neigh_build)" );
    cudaStream_t s = ctx->stream;
    std::vector<int> hc( n_total > 0 ? n_total : 1 );
    if ( n_total > 0 )
    {
        CBMD_CUDA( cudaMemcpyAsync( hc.data(), ctx->nb_pull ? ctx->nb_count_i : ctx->nb_count,
                                    (size_t)n_total * sizeof( int ), cudaMemcpyDeviceToHost, s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
        if ( ctx->nb_pull ) // ghost rows of a pull table hold only j-side entries
            std::fill( hc.begin() + n_local, hc.end(), 0 );
    }
    std::vector<int64_t> ho( n_local + 1, 0 );
    for ( int i = 0; i < n_local; i++ )
        ho[i + 1] = ho[i] + hc[i];
    if ( counts )
        std::copy( hc.begin(), hc.begin() + n_total, counts );
    if ( offsets )
        std::copy( ho.begin(), ho.end(), offsets );
    if ( neighbors && ho[n_local] > 0 )
    {
        const size_t ob = ( (size_t)( n_local + 1 ) * sizeof( int64_t ) + 255 ) & ~(size_t)255;
        char *st = (char *)cbmd_scratch( ctx, ob + (size_t)ho[n_local] * sizeof( int ) );
        int64_t *d_off = (int64_t *)st;
        int *d_csr = (int *)( st + ob );
        CBMD_CUDA( cudaMemcpyAsync( d_off, ho.data(), (size_t)( n_local + 1 ) * sizeof( int64_t ),
                                    cudaMemcpyHostToDevice, s ) );
        k_nb_to_csr<<<div_up( n_local, 256 ), 256, 0, s>>>( ctx->nb, ctx->nb_rows, ctx->nb_count, d_off,
                                                            n_local, d_csr );
        CBMD_LAUNCH_CHECK( ctx );
        CBMD_CUDA( cudaMemcpyAsync( neighbors, d_csr, (size_t)ho[n_local] * sizeof( int ),
                                    cudaMemcpyDeviceToHost, s ) );
        CBMD_CUDA( cudaStreamSynchronize( s ) );
    }
    CBMD_API_END
}

extern "C" int64_t cbmd_table_offset( int atom, int n, int row_capacity )
{
    return (int64_t)nb_entry( atom, n, ( row_capacity + 3 ) & ~3 );
}

extern "C" int64_t cbmd_table_size( int n_atoms, int row_capacity )
{
    return (int64_t)nb_table_size( ( n_atoms + 31 ) & ~31, ( row_capacity + 3 ) & ~3 );
}

// sum of the row lengths on the device (integer sum: order does not matter)
__global__ void __launch_bounds__( 256 )
    k_count_sum( const int *__restrict__ count, int n, unsigned long long *__restrict__ out )
{
    unsigned long long t = 0;
    for ( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
        t += (unsigned long long)count[i];
    for ( int o = 16; o > 0; o >>= 1 )
        t += __shfl_down_sync( 0xffffffffu, t, o );
    if ( ( threadIdx.x & 31 ) == 0 && t )
        atomicAdd( out, t );
}

extern "C" int cbmd_neigh_sizes( cbmd_ctx *ctx, int64_t *total, int *max_neigh )
{
    CBMD_API_BEGIN
    const int n_local = ctx->nb_n;
    if ( max_neigh )
        *max_neigh = ctx->nb_max;
    if ( total )
    {
        *total = 0;
        if ( n_local > 0 )
        {
            cudaStream_t s = ctx->stream;
            unsigned long long *d = (unsigned long long *)( ctx->d_flags + 28 ); // 8-byte aligned pair of ints
            CBMD_CUDA( cudaMemsetAsync( d, 0, sizeof( unsigned long long ), s ) );
            k_count_sum<<<std::min( div_up( n_local, 256 ), 1184 ), 256, 0, s>>>(
                ctx->nb_pull ? ctx->nb_count_i : ctx->nb_count, n_local, d );
            CBMD_LAUNCH_CHECK( ctx );
            unsigned long long *h = (unsigned long long *)( ctx->h_pinned_i + 12 );
            CBMD_CUDA( cudaMemcpyAsync( h, d, sizeof( unsigned long long ), cudaMemcpyDeviceToHost, s ) );
            CBMD_CUDA( cudaStreamSynchronize( s ) );
            *total = (int64_t)*h;
        }
    }
    CBMD_API_END
}
