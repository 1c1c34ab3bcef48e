// experiments/force_pairs.cu — "pair-shared rows": two spatially close atoms share ONE neighbour
// row (the union of their lists, 2 membership bits per entry).  The two lanes of a pair gather
// the same neighbour at the same step, so a warp gather touches 16 distinct sectors instead of
// ~27 and the index stream shrinks; the price is |union| > |own| pair evaluations per lane.
// NOT part of the product or the tests.
// Build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --extended-lambda -o experiments/force_pairs experiments/force_pairs.cu
#include "md_setup.h"

template <int R>
__device__ __forceinline__ double rcpd( double x )
{
    double r;
    asm( "rcp.approx.ftz.f64 %0, %1;" : "=d"( r ) : "d"( x ) );
    double e = fma( -x, r, 1.0 );
    if ( R == 2 )
        return fma( r, e, r );
    e = fma( e, e, e );
    r = fma( r, e, r );
    if ( R == 3 )
        return r;
    e = fma( -x, r, 1.0 );
    return fma( r, e, r );
}
__device__ __forceinline__ bool lt_pos( double a, double b ) { return __double_as_longlong( a ) < __double_as_longlong( b ); }

// pair table: ptile = 16 pairs (one warp); entry n of pair p at int4 index (ptile*rows4 + (n>>2))*16 + (p&15), component n&3
// entry = j | member0 << 30 | member1 << 31
#define IDX_MASK 0x3fffffff

__global__ void k_build_pairs( const XT *__restrict__ xt, const int *__restrict__ pair_atoms, int n_pairs,
                               const int *__restrict__ cell_start, const int *__restrict__ cell_atoms, int nc, double mn,
                               double rdx, double rsq, int *__restrict__ nbp, int rows4, int *__restrict__ pcnt,
                               int *__restrict__ own_cnt )
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if ( p >= n_pairs )
        return;
    const int a0 = pair_atoms[2 * p], a1 = pair_atoms[2 * p + 1];
    const XT x0 = xt[a0], x1 = xt[a1];
    int c0[3], c1[3];
    const double v0[3] = { x0.x, x0.y, x0.z }, v1[3] = { x1.x, x1.y, x1.z };
    for ( int d = 0; d < 3; d++ )
    {
        c0[d] = min( max( (int)floor( ( v0[d] - mn ) * rdx ), 0 ), nc - 1 );
        c1[d] = min( max( (int)floor( ( v1[d] - mn ) * rdx ), 0 ), nc - 1 );
    }
    // pairs live in one z column: c0[0]==c1[0], c0[1]==c1[1]
    const int zlo = max( min( c0[2], c1[2] ) - 1, 0 ), zhi = min( max( c0[2], c1[2] ) + 1, nc - 1 );
    int count = 0, n0 = 0, n1 = 0;
    int *base = nbp + ( (size_t)( p >> 4 ) * rows4 * 16 + ( p & 15 ) ) * 4;
    for ( int a = max( c0[0] - 1, 0 ); a <= min( c0[0] + 1, nc - 1 ); a++ )
        for ( int b = max( c0[1] - 1, 0 ); b <= min( c0[1] + 1, nc - 1 ); b++ )
        {
            const int row = ( a * nc + b ) * nc;
            const int s0 = cell_start[row + zlo], s1 = cell_start[row + zhi + 1];
            for ( int s = s0; s < s1; s++ )
            {
                const int j = cell_atoms[s];
                const XT xj = xt[j];
                double dx = x0.x - xj.x, dy = x0.y - xj.y, dz = x0.z - xj.z;
                const bool m0 = j != a0 && dx * dx + dy * dy + dz * dz <= rsq;
                dx = x1.x - xj.x, dy = x1.y - xj.y, dz = x1.z - xj.z;
                const bool m1 = a1 != a0 && j != a1 && dx * dx + dy * dy + dz * dz <= rsq;
                if ( m0 || m1 )
                {
                    if ( count < rows4 * 4 )
                        base[( count >> 2 ) * 64 + ( count & 3 )] = j | ( m0 ? 1 << 30 : 0 ) | ( m1 ? 1 << 31 : 0 );
                    count++;
                    n0 += m0;
                    n1 += m1;
                }
            }
        }
    for ( int k = count; k < ( ( count + 3 ) & ~3 ); k++ )
        base[( k >> 2 ) * 64 + ( k & 3 )] = a0; // no membership bits: never contributes
    pcnt[p] = count;
    atomicAdd( own_cnt, n0 + n1 );
}

struct ArgsP
{
    const int *pair_atoms;
    int n_pairs;
    const int4 *nbp;
    cudaTextureObject_t texnb;
    const int *pcnt;
    int rows4, cap;
    // fp64
    const XT *xt;
    const double2 *xy;
    cudaTextureObject_t texz;
    double *f;
    double lj1, lj2, cutsq;
    // fp32
    const float4 *xf;
    float *fs;
};

// FP32: lane l of a warp -> pair (warp*16 + l>>1), atom = pair_atoms[2*pair + (l&1)]
template <int U4, int IDXTEX, bool MEMBER>
__global__ void __launch_bounds__( 128 ) k_sp2( const __grid_constant__ ArgsP a )
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = g >> 1, side = g & 1;
    if ( p >= a.n_pairs )
        return;
    const int i = a.pair_atoms[2 * p + side];
    const bool dup = side == 1 && i == a.pair_atoms[2 * p]; // unpaired atom: second lane idles
    const float4 xi = __ldg( a.xf + i );
    float fx = 0, fy = 0, fz = 0;
    const int c4 = ( a.pcnt[p] + 3 ) >> 2;
    const int base = ( p >> 4 ) * a.rows4 * 16 + ( p & 15 );
    const float lj1 = (float)a.lj1, lj2 = (float)a.lj2, cutsq = (float)a.cutsq;
    const int mbit = 30 + side;
#pragma unroll( U4 )
    for ( int k4 = 0; k4 < c4; k4++ )
    {
        const int4 q = IDXTEX ? tex1Dfetch<int4>( a.texnb, base + k4 * 16 ) : __ldg( a.nbp + base + k4 * 16 );
        const int jj[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
        {
            const int e = jj[u], j = e & IDX_MASK;
            const float4 t = __ldg( a.xf + j );
            const float dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - t.z;
            const float rsq = dx * dx + dy * dy + dz * dz;
            const bool in = MEMBER ? ( rsq < cutsq && ( ( e >> mbit ) & 1 ) ) : ( rsq < cutsq && j != i );
            if ( in )
            {
                float r2inv;
                asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r2inv ) : "f"( rsq ) );
                const float r6inv = r2inv * r2inv * r2inv;
                const float fpair = ( r6inv * ( lj1 * r6inv - lj2 ) ) * r2inv;
                fx += dx * fpair;
                fy += dy * fpair;
                fz += dz * fpair;
            }
        }
    }
    if ( !dup )
    {
        a.fs[i] = fx;
        a.fs[(size_t)a.cap + i] = fy;
        a.fs[2 * (size_t)a.cap + i] = fz;
    }
}

template <int U4, int RCP, bool MEMBER>
__global__ void __launch_bounds__( 128 ) k_dp2( const __grid_constant__ ArgsP a )
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = g >> 1, side = g & 1;
    if ( p >= a.n_pairs )
        return;
    const int i = a.pair_atoms[2 * p + side];
    const bool dup = side == 1 && i == a.pair_atoms[2 * p];
    const XT xi = a.xt[i];
    double fx = 0, fy = 0, fz = 0;
    const int c4 = ( a.pcnt[p] + 3 ) >> 2;
    const int base = ( p >> 4 ) * a.rows4 * 16 + ( p & 15 );
    const double lj1 = a.lj1, lj2 = a.lj2, cutsq = a.cutsq;
    const int mbit = 30 + side;
#pragma unroll( U4 )
    for ( int k4 = 0; k4 < c4; k4++ )
    {
        const int4 q = __ldg( a.nbp + base + k4 * 16 );
        const int jj[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
        {
            const int e = jj[u], j = e & IDX_MASK;
            const double2 t = __ldg( a.xy + j );
            const int2 d = tex1Dfetch<int2>( a.texz, j );
            const double dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - __hiloint2double( d.y, d.x );
            const double rsq = dx * dx + dy * dy + dz * dz;
            const bool in = MEMBER ? ( lt_pos( rsq, cutsq ) && ( ( e >> mbit ) & 1 ) ) : ( lt_pos( rsq, cutsq ) && j != i );
            if ( in )
            {
                const double r2inv = rcpd<RCP>( rsq );
                const double r6inv = r2inv * r2inv * r2inv;
                const double fpair = ( r6inv * ( lj1 * r6inv - lj2 ) ) * r2inv;
                fx += dx * fpair;
                fy += dy * fpair;
                fz += dz * fpair;
            }
        }
    }
    if ( !dup )
    {
        a.f[i] = fx;
        a.f[(size_t)a.cap + i] = fy;
        a.f[2 * (size_t)a.cap + i] = fz;
    }
}

__global__ void __launch_bounds__( 128 )
    k_ref( const XT *__restrict__ xt, const int *__restrict__ nb, const int *__restrict__ cnt, int rows, int n,
           double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = xt[i];
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    for ( int k = 0; k < c; k++ )
    {
        const int j = nb[TB( i, rows ) + (size_t)k * 32];
        const XT xj = xt[j];
        const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        const double rsq = dx * dx + dy * dy + dz * dz;
        if ( rsq < cutsq )
        {
            const double r2inv = 1.0 / rsq;
            const double r6inv = r2inv * r2inv * r2inv;
            const double fpair = ( r6inv * ( lj1 * r6inv - lj2 ) ) * r2inv;
            fx += dx * fpair;
            fy += dy * fpair;
            fz += dz * fpair;
        }
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

int main( int argc, char **argv )
{
    const int cells = argc > 1 ? atoi( argv[1] ) : 100;
    const int reps = argc > 2 ? atoi( argv[2] ) : 5;
    const int mode = argc > 3 ? atoi( argv[3] ) : 2; // pairing: 0 index order, 1 morton, 2 greedy nearest
    MdSetup S;
    S.build( cells, 2.5, 0.3 );
    const int n = S.n, ntot = S.ntot, cap = S.cap;

    // ---- pairing of the owned atoms, per z column, cell by cell
    std::vector<int> pair_atoms;
    pair_atoms.reserve( n + 64 );
    {
        const int nc = S.nc;
        double sumd = 0;
        long long npaired = 0;
        std::vector<int> cur;
        for ( int col = 0; col < nc * nc; col++ )
        {
            int carry = -1;
            for ( int cz = 0; cz < nc; cz++ )
            {
                const int c = col * nc + cz;
                cur.clear();
                for ( int s = S.cell_start[c]; s < S.cell_start[c + 1]; s++ )
                    if ( S.cell_atoms[s] < n )
                        cur.push_back( S.cell_atoms[s] );
                if ( cur.empty() )
                    continue;
                auto dist2 = [&]( int a, int b )
                {
                    double d = 0;
                    for ( int k = 0; k < 3; k++ )
                    {
                        const double t = S.x[3 * (size_t)a + k] - S.x[3 * (size_t)b + k];
                        d += t * t;
                    }
                    return d;
                };
                if ( mode == 1 )
                {
                    auto morton = [&]( int a )
                    {
                        unsigned key = 0;
                        for ( int k = 0; k < 3; k++ )
                        {
                            const double cellw = 1.0 / S.rdx;
                            const double fr = ( S.x[3 * (size_t)a + k] - S.mn ) * S.rdx;
                            int b = (int)( ( fr - std::floor( fr ) ) * 4.0 );
                            b = std::min( std::max( b, 0 ), 3 );
                            (void)cellw;
                            for ( int bit = 0; bit < 2; bit++ )
                                key |= ( ( b >> bit ) & 1u ) << ( 3 * bit + ( 2 - k ) );
                        }
                        return key;
                    };
                    std::stable_sort( cur.begin(), cur.end(), [&]( int a, int b ) { return morton( a ) < morton( b ); } );
                }
                std::vector<char> used( cur.size(), 0 );
                auto nearest = [&]( int a )
                {
                    int best = -1;
                    double bd = 1e30;
                    for ( size_t k = 0; k < cur.size(); k++ )
                        if ( !used[k] && cur[k] != a )
                        {
                            const double d = dist2( a, cur[k] );
                            if ( d < bd )
                                bd = d, best = (int)k;
                        }
                    return best;
                };
                if ( carry >= 0 )
                {
                    const int k = mode == 2 ? nearest( carry ) : 0;
                    used[k] = 1;
                    pair_atoms.push_back( carry );
                    pair_atoms.push_back( cur[k] );
                    sumd += std::sqrt( dist2( carry, cur[k] ) );
                    npaired++;
                    carry = -1;
                }
                for ( size_t k = 0; k < cur.size(); k++ )
                {
                    if ( used[k] )
                        continue;
                    used[k] = 1;
                    int m = -1;
                    if ( mode == 2 )
                        m = nearest( cur[k] );
                    else
                        for ( size_t t = k + 1; t < cur.size(); t++ )
                            if ( !used[t] )
                            {
                                m = (int)t;
                                break;
                            }
                    if ( m < 0 )
                    {
                        carry = cur[k];
                        break;
                    }
                    used[m] = 1;
                    pair_atoms.push_back( cur[k] );
                    pair_atoms.push_back( cur[m] );
                    sumd += std::sqrt( dist2( cur[k], cur[m] ) );
                    npaired++;
                }
            }
            if ( carry >= 0 )
            {
                pair_atoms.push_back( carry );
                pair_atoms.push_back( carry );
            }
        }
        printf( "pairing mode %d: %zu pairs for %d atoms, mean partner distance %.3f\n", mode, pair_atoms.size() / 2, n,
                sumd / npaired );
    }
    const int n_pairs = (int)( pair_atoms.size() / 2 );
    const int np16 = ( n_pairs + 15 ) & ~15;

    std::vector<XT> hxt( cap );
    std::vector<double2> hxy( cap );
    std::vector<double> hz( cap );
    std::vector<float4> hxf( cap );
    for ( int i = 0; i < ntot; i++ )
    {
        const double *p = &S.x[3 * (size_t)i];
        hxt[i] = { p[0], p[1], p[2], 0 };
        hxy[i] = make_double2( p[0], p[1] );
        hz[i] = p[2];
        hxf[i] = make_float4( (float)p[0], (float)p[1], (float)p[2], 0.f );
    }
    std::vector<XT> hxt32( cap );
    for ( int i = 0; i < ntot; i++ )
        hxt32[i] = { (double)hxf[i].x, (double)hxf[i].y, (double)hxf[i].z, 0 };
    XT *d_xt, *d_xt32;
    double2 *d_xy;
    double *d_z, *d_f, *d_ref;
    float4 *d_xf;
    float *d_fs;
    int *d_cs, *d_ca, *d_nb, *d_cnt, *d_pa, *d_pcnt, *d_own;
    const int rows = 128, rows4 = 48, n32 = ( n + 31 ) & ~31;
    CK( cudaMalloc( &d_xt, (size_t)cap * sizeof( XT ) ) );
    CK( cudaMalloc( &d_xt32, (size_t)cap * sizeof( XT ) ) );
    CK( cudaMalloc( &d_xy, (size_t)cap * 24 ) );
    d_z = (double *)( d_xy + cap );
    CK( cudaMalloc( &d_xf, (size_t)cap * 16 ) );
    CK( cudaMalloc( &d_f, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_ref, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_fs, 3 * (size_t)cap * 4 ) );
    CK( cudaMalloc( &d_cs, S.cell_start.size() * 4 ) );
    CK( cudaMalloc( &d_ca, (size_t)ntot * 4 ) );
    CK( cudaMalloc( &d_nb, (size_t)rows * n32 * 4 ) );
    CK( cudaMalloc( &d_cnt, (size_t)cap * 4 ) );
    CK( cudaMalloc( &d_pa, pair_atoms.size() * 4 ) );
    CK( cudaMalloc( &d_pcnt, (size_t)np16 * 4 ) );
    CK( cudaMalloc( &d_own, 4 ) );
    int4 *d_nbp;
    CK( cudaMalloc( &d_nbp, (size_t)np16 * rows4 * 16 ) );
    CK( cudaMemcpy( d_xt, hxt.data(), (size_t)cap * sizeof( XT ), cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xt32, hxt32.data(), (size_t)cap * sizeof( XT ), cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xy, hxy.data(), (size_t)cap * 16, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_z, hz.data(), (size_t)cap * 8, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xf, hxf.data(), (size_t)cap * 16, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_cs, S.cell_start.data(), S.cell_start.size() * 4, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_ca, S.cell_atoms.data(), (size_t)ntot * 4, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_pa, pair_atoms.data(), pair_atoms.size() * 4, cudaMemcpyHostToDevice ) );
    CK( cudaMemset( d_own, 0, 4 ) );
    k_build_list<false><<<( n + 127 ) / 128, 128>>>( d_xt, n, d_cs, d_ca, S.nc, S.mn, S.rdx, S.rn * S.rn, d_nb, rows, d_cnt );
    k_build_pairs<<<( n_pairs + 127 ) / 128, 128>>>( d_xt, d_pa, n_pairs, d_cs, d_ca, S.nc, S.mn, S.rdx, S.rn * S.rn, (int *)d_nbp,
                                                     rows4, d_pcnt, d_own );
    CK( cudaDeviceSynchronize() );
    std::vector<int> hp( n_pairs ), hc( n );
    int own = 0;
    CK( cudaMemcpy( hp.data(), d_pcnt, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost ) );
    CK( cudaMemcpy( hc.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost ) );
    CK( cudaMemcpy( &own, d_own, 4, cudaMemcpyDeviceToHost ) );
    long long tot = 0, ptot = 0;
    int pmax = 0;
    for ( int c : hc )
        tot += c;
    for ( int c : hp )
        ptot += c, pmax = std::max( pmax, c );
    const double nn = (double)tot / n;
    printf( "own neighbours/atom %.2f (membership bits: %.2f); union row %.2f per pair (max %d, capacity %d)\n", nn,
            (double)own / n, (double)ptot / n_pairs, pmax, rows4 * 4 );

    auto make_tex = []( void *ptr, size_t bytes, cudaChannelFormatDesc desc )
    {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = ptr;
        rd.res.linear.desc = desc;
        rd.res.linear.sizeInBytes = bytes;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t t = 0;
        CK( cudaCreateTextureObject( &t, &rd, &td, nullptr ) );
        return t;
    };
    cudaTextureObject_t texz = make_tex( d_z, (size_t)cap * 8, cudaCreateChannelDesc<int2>() );
    cudaTextureObject_t texnb = make_tex( d_nbp, (size_t)np16 * rows4 * 16, cudaCreateChannelDesc<int4>() );

    const double lj1 = 48.0, lj2 = 24.0, cutsq = S.rc * S.rc;
    const int grid = ( n + 127 ) / 128;
    std::vector<double> ref( 3 * (size_t)cap ), ref32( 3 * (size_t)cap ), got( 3 * (size_t)cap );
    std::vector<float> gots( 3 * (size_t)cap );
    k_ref<<<grid, 128>>>( d_xt, d_nb, d_cnt, rows, n, d_ref, cap, lj1, lj2, cutsq );
    CK( cudaMemcpy( ref.data(), d_ref, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
    k_ref<<<grid, 128>>>( d_xt32, d_nb, d_cnt, rows, n, d_ref, cap, lj1, lj2, cutsq );
    CK( cudaMemcpy( ref32.data(), d_ref, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
    const double bytes64 = n * ( 4.0 * nn + 32.0 ) + (double)ntot * 28.0;
    const double bytes32 = n * ( 4.0 * nn + 12.0 + 8.0 ) + (double)ntot * 16.0;
    Timer T;
    ArgsP a = { d_pa, n_pairs, d_nbp, texnb, d_pcnt, rows4, cap, d_xt, d_xy, texz, d_f, lj1, lj2, cutsq, d_xf, d_fs };
    const int pgrid = ( 2 * n_pairs + 127 ) / 128;
    auto report = [&]( const char *name, float ms, double bytes, double err )
    {
        printf( "%-52s %8.4f ms  %7.1f GB/s alg (%.3f of 6547.5)  relerr %.2e\n", name, ms, bytes / ( ms * 1e-3 ) / 1e9,
                bytes / ( ms * 1e-3 ) / 1e9 / 6547.5, err );
        fflush( stdout );
    };
    auto run_d = [&]( const char *name, auto launch )
    {
        CK( cudaMemset( d_f, 0, 3 * (size_t)cap * 8 ) );
        const float ms = T.time( launch, reps );
        CK( cudaMemcpy( got.data(), d_f, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
        double mx = 0, mr = 0;
        for ( int c = 0; c < 3; c++ )
            for ( int i = 0; i < n; i++ )
            {
                mx = std::max( mx, std::fabs( got[(size_t)c * cap + i] - ref[(size_t)c * cap + i] ) );
                mr = std::max( mr, std::fabs( ref[(size_t)c * cap + i] ) );
            }
        report( name, ms, bytes64, mx / mr );
    };
    auto run_s = [&]( const char *name, auto launch )
    {
        CK( cudaMemset( d_fs, 0, 3 * (size_t)cap * 4 ) );
        const float ms = T.time( launch, reps );
        CK( cudaMemcpy( gots.data(), d_fs, 3 * (size_t)cap * 4, cudaMemcpyDeviceToHost ) );
        double mx = 0, mr = 0;
        for ( int c = 0; c < 3; c++ )
            for ( int i = 0; i < n; i++ )
            {
                mx = std::max( mx, std::fabs( (double)gots[(size_t)c * cap + i] - ref32[(size_t)c * cap + i] ) );
                mr = std::max( mr, std::fabs( ref32[(size_t)c * cap + i] ) );
            }
        report( name, ms, bytes32, mx / mr );
    };
    run_s( "p s fp32 pairs idx LDG u1x4 member", [&] { k_sp2<1, 0, true><<<pgrid, 128>>>( a ); } );
    run_s( "p s fp32 pairs idx LDG u2x4 member", [&] { k_sp2<2, 0, true><<<pgrid, 128>>>( a ); } );
    run_s( "p s fp32 pairs idx LDG u3x4 member", [&] { k_sp2<3, 0, true><<<pgrid, 128>>>( a ); } );
    run_s( "p s fp32 pairs idx TEX u2x4 member", [&] { k_sp2<2, 1, true><<<pgrid, 128>>>( a ); } );
    run_s( "p s fp32 pairs idx TEX u3x4 member", [&] { k_sp2<3, 1, true><<<pgrid, 128>>>( a ); } );
    run_s( "p s fp32 pairs idx LDG u2x4 nomember", [&] { k_sp2<2, 0, false><<<pgrid, 128>>>( a ); } );
    run_d( "p d fp64 pairs idx LDG u1x4 rcp3 member", [&] { k_dp2<1, 3, true><<<pgrid, 128>>>( a ); } );
    run_d( "p d fp64 pairs idx LDG u2x4 rcp3 member", [&] { k_dp2<2, 3, true><<<pgrid, 128>>>( a ); } );
    run_d( "p d fp64 pairs idx LDG u2x4 rcp3 nomember", [&] { k_dp2<2, 3, false><<<pgrid, 128>>>( a ); } );
    return 0;
}
