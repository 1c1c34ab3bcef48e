// experiments/force_variants.cu — standalone A/B harness for LJ force-kernel layouts.
// NOT part of the product or the tests: it exists to measure, on a real B200, how the
// position-gather layout changes L1/shared-memory wavefronts.  Build + run:
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp \
//        -o experiments/force_variants experiments/force_variants.cu
//   experiments/force_variants [cells=100] [reps=5]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#define CK( x )                                                                                   \
    do                                                                                            \
    {                                                                                             \
        cudaError_t e = ( x );                                                                    \
        if ( e != cudaSuccess )                                                                   \
        {                                                                                         \
            printf( "CUDA error %s at %s:%d\n", cudaGetErrorString( e ), __FILE__, __LINE__ );    \
            exit( 1 );                                                                            \
        }                                                                                         \
    } while ( 0 )

// the product's tiled Verlet table: neighbour k of atom i at nb[((i>>5)*rows + k)*32 + (i&31)];
// every kernel's `stride` argument is the row capacity of that table
#define TB( i, rows ) ( ( (size_t)( ( i ) >> 5 ) * (size_t)( rows ) ) * 32 + (size_t)( ( i ) & 31 ) )

struct alignas( 32 ) XT
{
    double x, y, z;
    long long t;
};

__device__ __forceinline__ XT ld_xt( const XT *p )
{
    XT r;
    double t;
    asm volatile( "ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                  : "=d"( r.x ), "=d"( r.y ), "=d"( r.z ), "=d"( t )
                  : "l"( p ) );
    r.t = __double_as_longlong( t );
    return r;
}

__device__ __forceinline__ double rcp5( double x )
{
    double r;
    asm( "rcp.approx.ftz.f64 %0, %1;" : "=d"( r ) : "d"( x ) );
    double e = fma( -x, r, 1.0 );
    e = fma( e, e, e );
    r = fma( r, e, r );
    e = fma( -x, r, 1.0 );
    r = fma( r, e, r );
    return r;
}
__device__ __forceinline__ double rcp3( double x )
{
    double r;
    asm( "rcp.approx.ftz.f64 %0, %1;" : "=d"( r ) : "d"( x ) );
    double e = fma( -x, r, 1.0 );
    e = fma( e, e, e );
    r = fma( r, e, r );
    return r;
}
// strict rsq < cutsq for non-negative doubles through the integer pipe
__device__ __forceinline__ bool lt_pos( double a, double b )
{
    return __double_as_longlong( a ) < __double_as_longlong( b );
}

#define LJ_BODY( RCP, CMP )                                                                       \
    const double rsq = dx * dx + dy * dy + dz * dz;                                               \
    if ( CMP )                                                                                    \
    {                                                                                             \
        const double r2inv = RCP( rsq );                                                          \
        const double r6inv = r2inv * r2inv * r2inv;                                               \
        const double fpair = ( r6inv * ( lj1 * r6inv - lj2 ) ) * r2inv;                           \
        fx += dx * fpair;                                                                         \
        fy += dy * fpair;                                                                         \
        fz += dz * fpair;                                                                         \
    }

// V0: product kernel as of round-1 first bench (AoS 32 B, LDG.256, rcp5, DSETP)
__global__ void __launch_bounds__( 128 )
    k_v0( const XT *__restrict__ xt, const int *__restrict__ nb, const int *__restrict__ cnt,
          int stride, int n, double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 4
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        const XT xj = ld_xt( xt + j );
        const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        LJ_BODY( rcp5, rsq < cutsq )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

// V0b: same gather, cheaper arithmetic (rcp3 + integer compare)
__global__ void __launch_bounds__( 128 )
    k_v0b( const XT *__restrict__ xt, const int *__restrict__ nb, const int *__restrict__ cnt,
           int stride, int n, double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 4
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        const XT xj = ld_xt( xt + j );
        const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        LJ_BODY( rcp3, lt_pos( rsq, cutsq ) )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

// V1: SoA positions, three LDG.64 gathers
template <int UNROLL>
__global__ void __launch_bounds__( 128 )
    k_v1( const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
          const int *__restrict__ nb, const int *__restrict__ cnt, int stride, int n,
          double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const double xi = x[i], yi = y[i], zi = z[i];
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll UNROLL
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        const double dx = xi - __ldg( x + j ), dy = yi - __ldg( y + j ), dz = zi - __ldg( z + j );
        LJ_BODY( rcp3, lt_pos( rsq, cutsq ) )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

// V2: AoS 32 B read as LDG.128 (x,y) + LDG.64 (z)
__global__ void __launch_bounds__( 128 )
    k_v2( const XT *__restrict__ xt, const int *__restrict__ nb, const int *__restrict__ cnt,
          int stride, int n, double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 4
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        const double2 xy = __ldg( (const double2 *)( xt + j ) );
        const double zz = __ldg( &xt[j].z );
        const double dx = xi.x - xy.x, dy = xi.y - xy.y, dz = xi.z - zz;
        LJ_BODY( rcp3, lt_pos( rsq, cutsq ) )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

// V3: AoS 24 B (packed xyz, no padding), three LDG.64 from the same line
__global__ void __launch_bounds__( 128 )
    k_v3( const double *__restrict__ x3, const int *__restrict__ nb, const int *__restrict__ cnt,
          int stride, int n, double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const double xi = x3[3 * (size_t)i], yi = x3[3 * (size_t)i + 1], zi = x3[3 * (size_t)i + 2];
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 4
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        const double *q = x3 + 3 * (size_t)j;
        const double dx = xi - __ldg( q ), dy = yi - __ldg( q + 1 ), dz = zi - __ldg( q + 2 );
        LJ_BODY( rcp3, lt_pos( rsq, cutsq ) )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}


// ---- texture-path gathers: the TEX front end of L1 instead of (or next to) the LSU pipe.
// MODE 0: every neighbour through two tex1Dfetch<int4>; MODE 1: odd neighbours through TEX,
// even ones through LDG.256; MODE 2: one in four through TEX.
__device__ __forceinline__ void tex_xt( cudaTextureObject_t tex, int j, double &x, double &y, double &z )
{
    const int4 a = tex1Dfetch<int4>( tex, 2 * j );
    const int4 b = tex1Dfetch<int4>( tex, 2 * j + 1 );
    x = __hiloint2double( a.y, a.x );
    y = __hiloint2double( a.w, a.z );
    z = __hiloint2double( b.y, b.x );
}
template <int MODE>
__global__ void __launch_bounds__( 128 )
    k_tex( const XT *__restrict__ xt, cudaTextureObject_t tex, const int *__restrict__ nb,
           const int *__restrict__ cnt, int stride, int n, double *__restrict__ f, int cap, double lj1,
           double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 4
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        double xj, yj, zj;
        const bool via_tex = MODE == 0 || ( MODE == 1 && ( k & 1 ) ) || ( MODE == 2 && ( k & 3 ) == 3 );
        if ( via_tex )
            tex_xt( tex, j, xj, yj, zj );
        else
        {
            const XT t = ld_xt( xt + j );
            xj = t.x;
            yj = t.y;
            zj = t.z;
        }
        const double dx = xi.x - xj, dy = xi.y - yj, dz = xi.z - zj;
        LJ_BODY( rcp5, rsq < cutsq )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}


// more TEX mixes.  FETCH 0: two int4 texels (32 B); 1: three int2 texels (24 B) of the same AoS
// record; 2: one int4 texel (x,y) through TEX + z through LDG.64 from the SoA z array.
// TEX_NUM of every TEX_DEN neighbours go through TEX, the rest through LDG.256.
template <int FETCH, int TEX_NUM, int TEX_DEN, int U>
__global__ void __launch_bounds__( 128 )
    k_tex2( const XT *__restrict__ xt, cudaTextureObject_t tex4, cudaTextureObject_t tex2,
            const double *__restrict__ zs, cudaTextureObject_t texz, const double2 *__restrict__ xy,
            const int *__restrict__ nb, const int *__restrict__ cnt,
            int stride, int n, double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll( U & 15 )
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        double xj, yj, zj;
        if ( ( k % TEX_DEN ) < TEX_NUM )
        {
            if ( FETCH == 7 )
            { // x through LDG.64 (SoA), y and z through two 8-byte TEX fetches
                const int2 b = tex1Dfetch<int2>( tex4, j ); // tex4 is bound to the SoA y array in this mode
                const int2 d = tex1Dfetch<int2>( texz, j );
                xj = __ldg( zs - 2 * (size_t)cap + j ); // d_soa holds x | y | z back to back
                yj = __hiloint2double( b.y, b.x );
                zj = __hiloint2double( d.y, d.x );
            }
            else if ( FETCH == 8 )
            { // no TEX at all: x,y through LDG.128 (packed), z through LDG.64 (SoA)
                const double2 t = __ldg( xy + j );
                xj = t.x;
                yj = t.y;
                zj = __ldg( zs + j );
            }
            else if ( FETCH == 5 || ( FETCH == 6 && ( k & 1 ) ) )
            { // x,y through TEX from the PACKED xy array (texel j), z through LDG.64
                const int4 a = tex1Dfetch<int4>( tex2, j ); // tex2 is bound to xy as int4 in this mode
                xj = __hiloint2double( a.y, a.x );
                yj = __hiloint2double( a.w, a.z );
                zj = __ldg( zs + j );
            }
            else if ( FETCH == 3 || FETCH == 6 )
            { // x,y through LDG.128 from the packed xy array, z through TEX (8 B texel)
                const double2 t = __ldg( xy + j );
                const int2 d = tex1Dfetch<int2>( texz, j );
                xj = t.x;
                yj = t.y;
                zj = __hiloint2double( d.y, d.x );
            }
            else if ( FETCH == 4 )
            { // x,y through TEX (16 B texel), z through TEX (8 B texel of the SoA z array)
                const int4 a = tex1Dfetch<int4>( tex4, 2 * j );
                const int2 d = tex1Dfetch<int2>( texz, j );
                xj = __hiloint2double( a.y, a.x );
                yj = __hiloint2double( a.w, a.z );
                zj = __hiloint2double( d.y, d.x );
            }
            else if ( FETCH == 0 )
                tex_xt( tex4, j, xj, yj, zj );
            else if ( FETCH == 1 )
            {
                const int2 a = tex1Dfetch<int2>( tex2, 4 * j );
                const int2 b = tex1Dfetch<int2>( tex2, 4 * j + 1 );
                const int2 d = tex1Dfetch<int2>( tex2, 4 * j + 2 );
                xj = __hiloint2double( a.y, a.x );
                yj = __hiloint2double( b.y, b.x );
                zj = __hiloint2double( d.y, d.x );
            }
            else
            {
                const int4 a = tex1Dfetch<int4>( tex4, 2 * j );
                xj = __hiloint2double( a.y, a.x );
                yj = __hiloint2double( a.w, a.z );
                zj = __ldg( zs + j );
            }
        }
        else
        {
            const XT t = ld_xt( xt + j );
            xj = t.x;
            yj = t.y;
            zj = t.z;
        }
        const double dx = xi.x - xj, dy = xi.y - yj, dz = xi.z - zj;
        LJ_BODY( rcp5, rsq < cutsq )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}


// ---- bank-aware row order.  ldg_patterns.cu shows that an LDG.256 warp gather is served one
// QUAD of lanes (4 x 32 B = the 128-byte data path) per cycle when the four sectors sit at
// four different positions of their 128-byte lines, whatever the lines are, and serialises on
// equal positions.  With 32-byte records the position is j & 3.  Reorder every row so that at
// step r lane q (= i & 3) reads a neighbour of class (r + q) & 3: a Latin square per quad, no
// conflict while all four classes of all four lanes last; the tail (classes run out unevenly)
// is filled with what is left.  The set of neighbours per row is unchanged.
template <int NC>
__global__ void __launch_bounds__( 128 )
    k_reorder_rows( int *__restrict__ nb, const int *__restrict__ cnt, int stride, int n, int mode )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const int c = cnt[i];
    int cls[NC][128 / NC + 16];
    int nc[NC], used[NC];
    for ( int t = 0; t < NC; t++ )
        nc[t] = used[t] = 0;
    for ( int k = 0; k < c; k++ )
    {
        const int j = nb[TB( i, stride ) + k * 32];
        const int b = j & ( NC - 1 );
        if ( nc[b] < 128 / NC + 16 )
            cls[b][nc[b]++] = j;
    }
    const int q = i & ( NC - 1 );
    for ( int r = 0; r < c; r++ )
    {
        int b = ( r + q ) & ( NC - 1 );
        if ( mode == 1 && used[b] >= nc[b] )
        {
            // class exhausted: take from the class with most entries left
            int best = -1, left = 0;
            for ( int t = 0; t < NC; t++ )
                if ( nc[t] - used[t] > left )
                {
                    left = nc[t] - used[t];
                    best = t;
                }
            b = best;
        }
        nb[TB( i, stride ) + r * 32] = cls[b][used[b]++];
    }
}


// the product's texture-assisted kernel shape with occupancy / unroll knobs
template <int U, int MINB, int BS>
__global__ void __launch_bounds__( BS, ( MINB % 100 ) )
    k_split( const XT *__restrict__ xt, const double2 *__restrict__ xy, cudaTextureObject_t texz,
             const int *__restrict__ nb, const int *__restrict__ cnt, int stride, int n,
             double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll( U )
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        double2 t;
        if ( MINB >= 100 )
            asm volatile( "ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"( t.x ), "=d"( t.y ) : "l"( xy + j ) );
        else
            t = __ldg( xy + j );
        const int2 d = tex1Dfetch<int2>( texz, j );
        const double dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - __hiloint2double( d.y, d.x );
        LJ_BODY( rcp5, rsq < cutsq )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

// DP-only bound: same arithmetic, j data synthesised in registers (no gather)
__global__ void __launch_bounds__( 128 )
    k_dp_only( const XT *__restrict__ xt, const int *__restrict__ nb, const int *__restrict__ cnt,
               int stride, int n, double *__restrict__ f, int cap, double lj1, double lj2,
               double cutsq )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = ld_xt( xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 4
    for ( int k = 0; k < c; k++ )
    {
        const int j = __ldg( p + k * 32 );
        const double s = (double)( j & 7 ) * 0.25;
        const double dx = 0.5 + s, dy = 0.25 - s, dz = 1.0 + 0.5 * s;
        LJ_BODY( rcp3, lt_pos( rsq, cutsq ) )
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
}

// index-stream-only bound
__global__ void __launch_bounds__( 128 )
    k_idx_only( const int *__restrict__ nb, const int *__restrict__ cnt, int stride, int n,
                double *__restrict__ f, int cap )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    int acc = 0;
    const int c = cnt[i];
    const int *p = nb + TB( i, stride );
#pragma unroll 8
    for ( int k = 0; k < c; k++ )
        acc += __ldg( p + k * 32 );
    f[i] = (double)acc;
}


// ---- shared-memory gather emulation: bank-conflict pattern of the real lists (j & (S-1)),
// staging cost of S atoms per CTA for BATCH*blockDim i-atoms.  Results are NOT physical.
template <int S, int BATCH, int MODE>
__global__ void __launch_bounds__( 256 )
    k_smem( const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
            const int *__restrict__ nb, const int *__restrict__ cnt, int stride, int n, int ntot,
            double *__restrict__ f, int cap, double lj1, double lj2, double cutsq )
{
    extern __shared__ double sm[];
    double *xs = sm, *ys = sm + S, *zs = sm + 2 * S;
    const int base = blockIdx.x * ( BATCH * 256 );
    for ( int k = threadIdx.x; k < S; k += 256 )
    {
        int g = base + k;
        g = g < ntot ? g : g - ntot;
        if ( MODE == 0 )
        {
            xs[k] = x[g];
            ys[k] = y[g];
            zs[k] = z[g];
        }
        else if ( MODE == 1 )
        { // xy interleaved (LDS.128) + z
            sm[2 * k] = x[g];
            sm[2 * k + 1] = y[g];
            zs[k] = z[g];
        }
        else
        { // 32-byte records
            sm[4 * k] = x[g];
            sm[4 * k + 1] = y[g];
            sm[4 * k + 2] = z[g];
        }
    }
    __syncthreads();
    for ( int b = 0; b < BATCH; b++ )
    {
        const int i = base + b * 256 + threadIdx.x;
        if ( i >= n )
            return;
        const double xi = x[i], yi = y[i], zi = z[i];
        double fx = 0, fy = 0, fz = 0;
        const int c = cnt[i];
        const int *p = nb + TB( i, stride );
#pragma unroll 4
        for ( int k = 0; k < c; k++ )
        {
            const int j = __ldg( p + k * 32 ) & ( S - 1 );
            double xj, yj, zj;
            if ( MODE == 0 )
            {
                xj = xs[j];
                yj = ys[j];
                zj = zs[j];
            }
            else if ( MODE == 1 )
            {
                const double2 t = *(const double2 *)( sm + 2 * j );
                xj = t.x;
                yj = t.y;
                zj = zs[j];
            }
            else
            {
                const double2 t = *(const double2 *)( sm + 4 * j );
                xj = t.x;
                yj = t.y;
                zj = sm[4 * j + 2];
            }
            const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
            LJ_BODY( rcp3, lt_pos( rsq, cutsq ) )
        }
        f[i] = fx;
        f[(size_t)cap + i] = fy;
        f[2 * (size_t)cap + i] = fz;
    }
}

// simple builder: thread per atom over the 27-cell stencil, ascending (cell, index)
__global__ void k_build( const XT *__restrict__ xt, int n, const int *__restrict__ cell_start,
                         const int *__restrict__ cell_atoms, int nc, double mn, double rdx,
                         double rsq, int *__restrict__ nb, int stride, int rows,
                         int *__restrict__ cnt )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = xt[i];
    int c[3];
    const double v[3] = { xi.x, xi.y, xi.z };
    for ( int d = 0; d < 3; d++ )
    {
        int q = (int)floor( ( v[d] - mn ) * rdx );
        c[d] = min( max( q, 0 ), nc - 1 );
    }
    int count = 0;
    for ( int a = max( c[0] - 1, 0 ); a <= min( c[0] + 1, nc - 1 ); a++ )
        for ( int b = max( c[1] - 1, 0 ); b <= min( c[1] + 1, nc - 1 ); b++ )
        {
            const int row = ( a * nc + b ) * nc;
            const int s0 = cell_start[row + max( c[2] - 1, 0 )],
                      s1 = cell_start[row + min( c[2] + 1, nc - 1 ) + 1];
            for ( int s = s0; s < s1; s++ )
            {
                const int j = cell_atoms[s];
                const XT xj = xt[j];
                const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                if ( j != i && dx * dx + dy * dy + dz * dz <= rsq )
                {
                    if ( count < rows )
                        nb[TB( i, stride ) + count * 32] = j;
                    count++;
                }
            }
        }
    cnt[i] = count;
}

int main( int argc, char **argv )
{
    const int cells = argc > 1 ? atoi( argv[1] ) : 100;
    const int reps = argc > 2 ? atoi( argv[2] ) : 5;
    const double a = std::cbrt( 4.0 / 0.8442 ), L = a * cells, rn = 2.8, rc = 2.5;
    const int n = 4 * cells * cells * cells;
    printf( "cells %d  atoms %d  L %.3f\n", cells, n, L );
    std::vector<double> x( 3 * (size_t)n );
    {
        const double basis[4][3] = { { 0, 0, 0 }, { .5, .5, 0 }, { .5, 0, .5 }, { 0, .5, .5 } };
        std::mt19937_64 rng( 12345 );
        std::uniform_real_distribution<double> u( -0.22, 0.22 );
        size_t k = 0;
        for ( int iz = 0; iz < cells; iz++ )
            for ( int iy = 0; iy < cells; iy++ )
                for ( int ix = 0; ix < cells; ix++ )
                    for ( int b = 0; b < 4; b++ )
                    {
                        const int ii[3] = { ix, iy, iz };
                        for ( int d = 0; d < 3; d++ )
                        {
                            double v = a * ( ii[d] + basis[b][d] ) + u( rng );
                            if ( v < 0 )
                                v += L;
                            if ( v >= L )
                                v -= L;
                            x[3 * k + d] = v;
                        }
                        k++;
                    }
    }
    // cell grid as the product derives it
    const int nbin = (int)( L / rn );
    const double dbin = L / nbin, eps = dbin / 1000;
    const double mn = -dbin - eps, mx = L + dbin + eps;
    const int nc = (int)std::floor( ( mx - mn ) / dbin );
    const double rdx = 1.0 / ( ( mx - mn ) / nc );
    auto cell_of = [&]( const double *p )
    {
        int c[3];
        for ( int d = 0; d < 3; d++ )
        {
            int q = (int)std::floor( ( p[d] - mn ) * rdx );
            c[d] = std::min( std::max( q, 0 ), nc - 1 );
        }
        return ( c[0] * nc + c[1] ) * nc + c[2];
    };
    // sort locals by cell (stable)
    {
        std::vector<int> cell( n ), order( n );
        for ( int i = 0; i < n; i++ )
            cell[i] = cell_of( &x[3 * (size_t)i] );
        std::iota( order.begin(), order.end(), 0 );
        std::stable_sort( order.begin(), order.end(),
                          [&]( int p, int q ) { return cell[p] < cell[q]; } );
        std::vector<double> y( x.size() );
        for ( int i = 0; i < n; i++ )
            for ( int d = 0; d < 3; d++ )
                y[3 * (size_t)i + d] = x[3 * (size_t)order[i] + d];
        x.swap( y );
    }
    // ghosts: 6 phases
    for ( int ph = 0; ph < 6; ph++ )
    {
        const int d = ph / 2;
        const size_t cur = x.size() / 3;
        static size_t last_recv = 0;
        const size_t np = cur - ( ph % 2 ? last_recv : 0 );
        size_t added = 0;
        for ( size_t i = 0; i < np; i++ )
        {
            const double c = x[3 * i + d];
            const bool sel = ( ph % 2 == 0 ) ? ( c >= L - rn ) : ( c <= rn );
            if ( sel )
            {
                double p[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
                p[d] += ( ph % 2 == 0 ) ? -L : L;
                x.insert( x.end(), p, p + 3 );
                added++;
            }
        }
        last_recv = added;
    }
    const int ntot = (int)( x.size() / 3 );
    printf( "ghosts %d (%.1f%%)  cells/dim %d\n", ntot - n, 100.0 * ( ntot - n ) / n, nc );
    // cell lists over all atoms
    const int ncells = nc * nc * nc;
    std::vector<int> cell_start( ncells + 1, 0 ), cell_atoms( ntot ), acell( ntot );
    for ( int i = 0; i < ntot; i++ )
    {
        acell[i] = cell_of( &x[3 * (size_t)i] );
        cell_start[acell[i] + 1]++;
    }
    for ( int c = 0; c < ncells; c++ )
        cell_start[c + 1] += cell_start[c];
    {
        std::vector<int> cur( cell_start.begin(), cell_start.end() - 1 );
        for ( int i = 0; i < ntot; i++ )
            cell_atoms[cur[acell[i]]++] = i;
    }
    // device data
    const int cap = ( ntot + 127 ) & ~127;
    std::vector<XT> hxt( cap );
    std::vector<double> hsoa( 3 * (size_t)cap, 0.0 );
    for ( int i = 0; i < ntot; i++ )
    {
        hxt[i] = { x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], 0 };
        for ( int d = 0; d < 3; d++ )
            hsoa[(size_t)d * cap + i] = x[3 * (size_t)i + d];
    }
    XT *d_xt;
    double *d_soa, *d_x3, *d_f, *d_f0;
    int *d_cs, *d_ca, *d_nb, *d_cnt;
    const int rows = 128, stride = rows, natoms32 = ( n + 31 ) & ~31;
    CK( cudaMalloc( &d_xt, (size_t)cap * sizeof( XT ) ) );
    CK( cudaMalloc( &d_soa, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_x3, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_f, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_f0, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_cs, ( ncells + 1 ) * 4 ) );
    CK( cudaMalloc( &d_ca, (size_t)ntot * 4 ) );
    CK( cudaMalloc( &d_nb, (size_t)rows * natoms32 * 4 ) );
    CK( cudaMalloc( &d_cnt, (size_t)cap * 4 ) );
    CK( cudaMemcpy( d_xt, hxt.data(), (size_t)cap * sizeof( XT ), cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_soa, hsoa.data(), 3 * (size_t)cap * 8, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_x3, x.data(), x.size() * 8, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_cs, cell_start.data(), ( ncells + 1 ) * 4, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_ca, cell_atoms.data(), (size_t)ntot * 4, cudaMemcpyHostToDevice ) );
    k_build<<<( n + 127 ) / 128, 128>>>( d_xt, n, d_cs, d_ca, nc, mn, rdx, rn * rn, d_nb, stride,
                                         rows, d_cnt );
    CK( cudaDeviceSynchronize() );
    std::vector<int> hcnt( n );
    CK( cudaMemcpy( hcnt.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost ) );
    long long tot = 0;
    int mxc = 0;
    for ( int c : hcnt )
    {
        tot += c;
        mxc = std::max( mxc, c );
    }
    const double nn = (double)tot / n;
    printf( "neighbours/atom %.2f  max %d\n", nn, mxc );
    const double bytes = n * ( 4.0 * nn + 32.0 ) + (double)ntot * 28.0;
    const double lj1 = 48.0, lj2 = 24.0, cutsq = rc * rc;
    cudaEvent_t e0, e1;
    cudaEventCreate( &e0 );
    cudaEventCreate( &e1 );
    const int grid = ( n + 127 ) / 128;
    double *x_ = d_soa, *y_ = d_soa + cap, *z_ = d_soa + 2 * (size_t)cap;
    std::vector<double> ref( 3 * (size_t)cap ), got( 3 * (size_t)cap );
    auto run = [&]( const char *name, auto launch, bool check )
    {
        CK( cudaMemset( d_f, 0, 3 * (size_t)cap * 8 ) );
        launch();
        launch();
        CK( cudaDeviceSynchronize() );
        cudaEventRecord( e0 );
        for ( int r = 0; r < reps; r++ )
            launch();
        cudaEventRecord( e1 );
        CK( cudaDeviceSynchronize() );
        float ms;
        cudaEventElapsedTime( &ms, e0, e1 );
        ms /= reps;
        double err = -1;
        if ( check )
        {
            CK( cudaMemcpy( got.data(), d_f, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
            double mx = 0, mr = 0;
            for ( int c = 0; c < 3; c++ )
                for ( int i = 0; i < n; i++ )
                {
                    mx = std::max( mx, std::fabs( got[(size_t)c * cap + i] - ref[(size_t)c * cap + i] ) );
                    mr = std::max( mr, std::fabs( ref[(size_t)c * cap + i] ) );
                }
            err = mx / mr;
        }
        printf( "%-28s %8.4f ms  %7.1f GB/s algorithmic (%.1f%% of 6554)  relerr %.2e\n", name, ms,
                bytes / ( ms * 1e-3 ) / 1e9, 100 * bytes / ( ms * 1e-3 ) / 1e9 / 6554, err );
    };
    // reference result
    k_v0<<<grid, 128>>>( d_xt, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq );
    CK( cudaDeviceSynchronize() );
    CK( cudaMemcpy( ref.data(), d_f, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
    run( "v0 AoS32 LDG.256 rcp5", [&] { k_v0<<<grid, 128>>>( d_xt, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    run( "v0b AoS32 LDG.256 rcp3+icmp", [&] { k_v0b<<<grid, 128>>>( d_xt, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    run( "v1 SoA 3xLDG.64 u4", [&] { k_v1<4><<<grid, 128>>>( x_, y_, z_, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    run( "v1 SoA 3xLDG.64 u8", [&] { k_v1<8><<<grid, 128>>>( x_, y_, z_, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    run( "v1 SoA 3xLDG.64 u2", [&] { k_v1<2><<<grid, 128>>>( x_, y_, z_, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    run( "v2 AoS32 LDG.128+LDG.64", [&] { k_v2<<<grid, 128>>>( d_xt, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    run( "v3 AoS24 3xLDG.64", [&] { k_v3<<<grid, 128>>>( d_x3, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = d_xt;
        rd.res.linear.desc = cudaCreateChannelDesc<int4>();
        rd.res.linear.sizeInBytes = (size_t)cap * sizeof( XT );
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        CK( cudaCreateTextureObject( &tex, &rd, &td, nullptr ) );
        run( "tex: all via 2x tex1Dfetch", [&] { k_tex<0><<<grid, 128>>>( d_xt, tex, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
        run( "tex: 1/2 TEX + 1/2 LDG.256", [&] { k_tex<1><<<grid, 128>>>( d_xt, tex, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
        cudaResourceDesc rd2 = rd;
        rd2.res.linear.desc = cudaCreateChannelDesc<int2>();
        cudaTextureObject_t tex2 = 0;
        CK( cudaCreateTextureObject( &tex2, &rd2, &td, nullptr ) );
        cudaResourceDesc rdz = rd;
        rdz.res.linear.devPtr = z_;
        rdz.res.linear.desc = cudaCreateChannelDesc<int2>();
        rdz.res.linear.sizeInBytes = (size_t)cap * 8;
        cudaTextureObject_t texz = 0;
        CK( cudaCreateTextureObject( &texz, &rdz, &td, nullptr ) );
        std::vector<double2> hxy( cap );
        for ( int i = 0; i < ntot; i++ )
            hxy[i] = make_double2( x[3 * (size_t)i], x[3 * (size_t)i + 1] );
        double2 *d_xy;
        CK( cudaMalloc( &d_xy, (size_t)cap * 16 ) );
        CK( cudaMemcpy( d_xy, hxy.data(), (size_t)cap * 16, cudaMemcpyHostToDevice ) );
#define RUN_TEX2( NAME, F, A, B, U )                                                              \
    run( NAME, [&] { k_tex2<F, A, B, U><<<grid, 128>>>( d_xt, tex, tex2, z_, texz, d_xy, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true )
        {
            cudaResourceDesc rdxy = rd;
            rdxy.res.linear.devPtr = d_xy;
            rdxy.res.linear.desc = cudaCreateChannelDesc<int4>();
            rdxy.res.linear.sizeInBytes = (size_t)cap * 16;
            cudaTextureObject_t texxy = 0;
            CK( cudaCreateTextureObject( &texxy, &rdxy, &td, nullptr ) );
#define RUN_TEX3( NAME, F, A, B, U )                                                              \
    run( NAME, [&] { k_tex2<F, A, B, U><<<grid, 128>>>( d_xt, tex, texxy, z_, texz, d_xy, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true )
            RUN_TEX3( "tex3 xy-texP+z-ldg64 all u4", 5, 1, 1, 4 );
            RUN_TEX3( "tex3 xy-texP+z-ldg64 all u8", 5, 1, 1, 8 );
            RUN_TEX3( "tex3 alternate roles u4", 6, 1, 1, 4 );
            RUN_TEX3( "tex3 alternate roles u8", 6, 1, 1, 8 );
            RUN_TEX3( "tex3 xy-texP+z-ldg64 3/4 u4", 5, 3, 4, 4 );
        }
        {
            cudaResourceDesc rdy = rdz;
            rdy.res.linear.devPtr = y_;
            cudaTextureObject_t texy = 0;
            CK( cudaCreateTextureObject( &texy, &rdy, &td, nullptr ) );
            run( "x-ldg64 + y-tex8 + z-tex8 u6", [&] { k_tex2<7, 1, 1, 6><<<grid, 128>>>( d_xt, texy, tex2, z_, texz, d_xy, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
            run( "x-ldg64 + y-tex8 + z-tex8 u4", [&] { k_tex2<7, 1, 1, 4><<<grid, 128>>>( d_xt, texy, tex2, z_, texz, d_xy, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
        }
        RUN_TEX2( "no TEX: xy-ldg128 + z-ldg64 u6", 8, 1, 1, 6 );
        RUN_TEX2( "no TEX: xy-ldg128 + z-ldg64 u4", 8, 1, 1, 4 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex all u8", 3, 1, 1, 8 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex all u4 rcp3", 3, 1, 1, 16 + 4 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex all u8 rcp3", 3, 1, 1, 16 + 8 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex all u6", 3, 1, 1, 6 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex all u2", 3, 1, 1, 2 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex 7/8 u8", 3, 7, 8, 8 );
        RUN_TEX2( "tex2 int4x2  1/2 u4", 0, 1, 2, 4 );
        RUN_TEX2( "tex2 int4x2  1/2 u8", 0, 1, 2, 8 );
        RUN_TEX2( "tex2 int4x2  3/8 u8", 0, 3, 8, 8 );
        RUN_TEX2( "tex2 int4x2  5/8 u8", 0, 5, 8, 8 );
        RUN_TEX2( "tex2 xy-tex+z-ldg64 all u4", 2, 1, 1, 4 );
        RUN_TEX2( "tex2 xy-tex+z-ldg64 3/4 u4", 2, 3, 4, 4 );
        RUN_TEX2( "tex2 xy-tex+z-ldg64 3/4 u8", 2, 3, 4, 8 );
        RUN_TEX2( "tex2 xy-tex+z-ldg64 7/8 u8", 2, 7, 8, 8 );
        RUN_TEX2( "tex2 xy-tex+z-ldg64 5/8 u8", 2, 5, 8, 8 );
        RUN_TEX2( "tex2 xy-tex+z-ldg64 1/2 u4", 2, 1, 2, 4 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex all u4", 3, 1, 1, 4 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex 3/4 u4", 3, 3, 4, 4 );
        RUN_TEX2( "tex2 xy-ldg128+z-tex 1/2 u4", 3, 1, 2, 4 );
        RUN_TEX2( "tex2 xy-tex+z-tex all u4", 4, 1, 1, 4 );
        RUN_TEX2( "tex2 xy-tex+z-tex 1/2 u4", 4, 1, 2, 4 );
        RUN_TEX2( "tex2 xy-tex+z-tex 3/4 u4", 4, 3, 4, 4 );
        run( "tex: 1/4 TEX + 3/4 LDG.256", [&] { k_tex<2><<<grid, 128>>>( d_xt, tex, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
    }
    run( "dp-only (no gather)", [&] { k_dp_only<<<grid, 128>>>( d_xt, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, false );
    run( "index-stream only", [&] { k_idx_only<<<grid, 128>>>( d_nb, d_cnt, stride, n, d_f, cap ); }, false );

    {
        // bank-aware row orders (must come last: they permute the table in place)
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = z_;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>();
        rd.res.linear.sizeInBytes = (size_t)cap * 8;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t texz = 0;
        CK( cudaCreateTextureObject( &texz, &rd, &td, nullptr ) );
        std::vector<double2> hxy( cap );
        for ( int i = 0; i < ntot; i++ )
            hxy[i] = make_double2( x[3 * (size_t)i], x[3 * (size_t)i + 1] );
        double2 *d_xy;
        CK( cudaMalloc( &d_xy, (size_t)cap * 16 ) );
        CK( cudaMemcpy( d_xy, hxy.data(), (size_t)cap * 16, cudaMemcpyHostToDevice ) );
        auto both = [&]( const char *tag )
        {
            char name[96];
            snprintf( name, sizeof name, "v0 LDG.256, rows %s", tag );
            run( name, [&] { k_v0<<<grid, 128>>>( d_xt, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
            snprintf( name, sizeof name, "xy-ldg128+z-tex u4, rows %s", tag );
            run( name, [&] { k_tex2<3, 1, 1, 4><<<grid, 128>>>( d_xt, 0, 0, z_, texz, d_xy, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
            snprintf( name, sizeof name, "xy-ldg128+z-tex u8, rows %s", tag );
            run( name, [&] { k_tex2<3, 1, 1, 8><<<grid, 128>>>( d_xt, 0, 0, z_, texz, d_xy, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true );
        };
#define RUN_SPLIT( U, MINB, BS )                                                                  \
    run( "split U" #U " minb" #MINB " bs" #BS, [&] { k_split<U, MINB, BS><<<( n + BS - 1 ) / BS, BS>>>( d_xt, d_xy, texz, d_nb, d_cnt, stride, n, d_f, cap, lj1, lj2, cutsq ); }, true )
        RUN_SPLIT( 4, 1, 128 );
        RUN_SPLIT( 6, 1, 128 );
        RUN_SPLIT( 4, 10, 128 );
        RUN_SPLIT( 6, 10, 128 );
        RUN_SPLIT( 4, 12, 128 );
        RUN_SPLIT( 6, 12, 128 );
        RUN_SPLIT( 8, 12, 128 );
        RUN_SPLIT( 4, 101, 128 );
        RUN_SPLIT( 6, 101, 128 );
        RUN_SPLIT( 6, 5, 256 );
        RUN_SPLIT( 6, 20, 64 );
        k_reorder_rows<8><<<grid, 128>>>( d_nb, d_cnt, stride, n, 1 );
        CK( cudaDeviceSynchronize() );
        both( "8-class" );
        k_reorder_rows<4><<<grid, 128>>>( d_nb, d_cnt, stride, n, 1 );
        CK( cudaDeviceSynchronize() );
        both( "4-class" );
        k_reorder_rows<2><<<grid, 128>>>( d_nb, d_cnt, stride, n, 1 );
        CK( cudaDeviceSynchronize() );
        both( "2-class" );
    }
    {
        constexpr int S = 4096, B = 4;
        const int g2 = ( n + B * 256 - 1 ) / ( B * 256 );
        cudaFuncSetAttribute( k_smem<S, B, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * S * 8 );
        cudaFuncSetAttribute( k_smem<S, B, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * S * 8 );
        cudaFuncSetAttribute( k_smem<S, B, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * S * 8 );
        run( "smem SoA 3xLDS.64 (emul)", [&] { k_smem<S, B, 0><<<g2, 256, 3 * S * 8>>>( x_, y_, z_, d_nb, d_cnt, stride, n, ntot, d_f, cap, lj1, lj2, cutsq ); }, false );
        run( "smem xy LDS.128 + z (emul)", [&] { k_smem<S, B, 1><<<g2, 256, 3 * S * 8>>>( x_, y_, z_, d_nb, d_cnt, stride, n, ntot, d_f, cap, lj1, lj2, cutsq ); }, false );
        run( "smem AoS32 LDS.128+64 (emul)", [&] { k_smem<S, B, 2><<<g2, 256, 4 * S * 8>>>( x_, y_, z_, d_nb, d_cnt, stride, n, ntot, d_f, cap, lj1, lj2, cutsq ); }, false );
    }
    return 0;
}
