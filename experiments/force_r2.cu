// experiments/force_r2.cu — round-2 A/B harness for the LJ full-list force kernel.
// NOT part of the product or the tests.  Legs:
//   d*  FP64: product kernel (xy LDG.128 + z TEX) with cheaper reciprocal / integer compare,
//       SM-local persistent tile scheduling (natural and blob order), index-stream cache hints,
//       int4 index loads
//   s*  FP32: float4 records through LDG.128 / TEX / both, float2 + z split, persistent
// Build:  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o experiments/force_r2 experiments/force_r2.cu
// Run:    experiments/force_r2 [cells=100] [reps=5] [name-filter]
#include "md_setup.h"

// ----------------------------------------------------------------------------- FP64 pieces
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
__device__ __forceinline__ bool lt_pos( double a, double b )
{
    return __double_as_longlong( a ) < __double_as_longlong( b );
}
__device__ __forceinline__ XT ld_xt( const XT *p )
{
    XT r;
    double t;
    asm volatile( "ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"( r.x ), "=d"( r.y ), "=d"( r.z ), "=d"( t ) : "l"( p ) );
    r.t = __double_as_longlong( t );
    return r;
}
__device__ __forceinline__ unsigned long long policy_evict_first()
{
    unsigned long long pol;
    asm( "createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"( pol ) );
    return pol;
}
__device__ __forceinline__ unsigned long long policy_evict_last()
{
    unsigned long long pol;
    asm( "createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"( pol ) );
    return pol;
}
__device__ __forceinline__ int ld_idx_pol( const int *p, unsigned long long pol )
{
    int v;
    asm volatile( "ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"( v ) : "l"( p ), "l"( pol ) );
    return v;
}
__device__ __forceinline__ double2 ld_xy_pol( const double2 *p, unsigned long long pol )
{
    double2 v;
    asm volatile( "ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"( v.x ), "=d"( v.y ) : "l"( p ), "l"( pol ) );
    return v;
}
__device__ __forceinline__ float4 ld_f4_pol( const float4 *p, unsigned long long pol )
{
    float4 v;
    asm volatile( "ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                  : "=f"( v.x ), "=f"( v.y ), "=f"( v.z ), "=f"( v.w )
                  : "l"( p ), "l"( pol ) );
    return v;
}
template <int H>
__device__ __forceinline__ int ld_idx( const int *p )
{
    int v;
    if ( H == 1 )
        asm volatile( "ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"( v ) : "l"( p ) );
    else if ( H == 2 )
        asm volatile( "ld.global.cs.b32 %0, [%1];" : "=r"( v ) : "l"( p ) );
    else
        v = __ldg( p );
    return v;
}

struct ArgsD
{
    const XT *xt;
    const double2 *xy;
    cudaTextureObject_t texz;
    const int *nb;
    const int *cnt;
    int rows, n;
    double *f;
    int cap;
    double lj1, lj2, cutsq;
};

// one lane = one atom; i = tile*32 + lane
template <int U, int RCP, bool ICMP, int IDXH>
__device__ __forceinline__ void atom_d( const ArgsD &a, int i )
{
    if ( i >= a.n )
        return;
    const XT xi = ld_xt( a.xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = a.cnt[i];
    const int *p = a.nb + TB( i, a.rows );
    const double lj1 = a.lj1, lj2 = a.lj2, cutsq = a.cutsq;
    const unsigned long long pf = policy_evict_first(), pl = policy_evict_last();
#pragma unroll( U )
    for ( int k = 0; k < c; k++ )
    {
        constexpr int H0 = IDXH < 3 ? IDXH : 0;
        const int j = IDXH >= 3 ? ld_idx_pol( p + k * 32, pf ) : ld_idx<H0>( p + k * 32 );
        const double2 t = IDXH == 4 ? ld_xy_pol( a.xy + j, pl ) : __ldg( a.xy + j );
        const int2 d = tex1Dfetch<int2>( a.texz, j );
        const double dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - __hiloint2double( d.y, d.x );
        const double rsq = dx * dx + dy * dy + dz * dz;
        if ( ICMP ? lt_pos( rsq, cutsq ) : ( rsq < cutsq ) )
        {
            const double r2inv = rcpd<RCP>( rsq );
            const double r6inv = r2inv * r2inv * r2inv;
            const double fpair = ( r6inv * ( lj1 * r6inv - lj2 ) ) * r2inv;
            fx += dx * fpair;
            fy += dy * fpair;
            fz += dz * fpair;
        }
    }
    a.f[i] = fx;
    a.f[(size_t)a.cap + i] = fy;
    a.f[2 * (size_t)a.cap + i] = fz;
}

template <int U, int RCP, bool ICMP, int IDXH>
__global__ void __launch_bounds__( 128 ) k_d( const __grid_constant__ ArgsD a )
{
    atom_d<U, RCP, ICMP, IDXH>( a, blockIdx.x * blockDim.x + threadIdx.x );
}

// SM-local persistent scheduling: the tiles (32 consecutive atoms) are cut into nsm contiguous
// ranges; the warps resident on SM s pull tiles from range s through one counter, so the
// warps that share an L1 work on neighbouring atoms.  A warp whose range is exhausted steals
// from the following ranges (completion does not depend on CTA placement).
template <class Body>
__device__ __forceinline__ void persistent_tiles( int *counters, int nsm, int n_tiles, const int *order, Body body )
{
    unsigned smid;
    asm volatile( "mov.u32 %0, %%smid;" : "=r"( smid ) );
    const int lane = threadIdx.x & 31;
    int s = (int)( smid % (unsigned)nsm );
    for ( int hop = 0; hop < nsm; hop++ )
    {
        const int lo = (int)( (long long)s * n_tiles / nsm ), hi = (int)( (long long)( s + 1 ) * n_tiles / nsm );
        if ( *( (volatile int *)counters + s ) < hi - lo )
            for ( ;; )
            {
                int t = 0;
                if ( lane == 0 )
                    t = atomicAdd( counters + s, 1 );
                t = __shfl_sync( 0xffffffffu, t, 0 ) + lo;
                if ( t >= hi )
                    break;
                body( order ? order[t] : t );
            }
        s = s + 1 == nsm ? 0 : s + 1;
    }
}

template <int U, int RCP, bool ICMP, int IDXH>
__global__ void __launch_bounds__( 128 )
    k_dp( const __grid_constant__ ArgsD a, int *counters, int nsm, int n_tiles, const int *order )
{
    const int lane = threadIdx.x & 31;
    persistent_tiles( counters, nsm, n_tiles, order, [&]( int t ) { atom_d<U, RCP, ICMP, IDXH>( a, t * 32 + lane ); } );
}

// int4 index loads: table nb4[((tile*rows4 + k4)*32 + lane)] holds entries 4*k4..4*k4+3 of the lane's row
template <int U4, int RCP>
__device__ __forceinline__ void atom_d4( const ArgsD &a, const int4 *nb4, int rows4, int i )
{
    if ( i >= a.n )
        return;
    const XT xi = ld_xt( a.xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = a.cnt[i];
    const int4 *p = nb4 + ( (size_t)( i >> 5 ) * rows4 ) * 32 + ( i & 31 );
    const double lj1 = a.lj1, lj2 = a.lj2, cutsq = a.cutsq;
    const int c4 = ( c + 3 ) >> 2;
#pragma unroll( U4 )
    for ( int k4 = 0; k4 < c4; k4++ )
    {
        const int4 q = __ldg( p + k4 * 32 );
        const int jj[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
        {
            const int j = jj[u]; // padded with the atom itself (rsq = 0 is rejected below)
            const double2 t = __ldg( a.xy + j );
            const int2 d = tex1Dfetch<int2>( a.texz, j );
            const double dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - __hiloint2double( d.y, d.x );
            const double rsq = dx * dx + dy * dy + dz * dz;
            if ( lt_pos( rsq, cutsq ) && j != i )
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
    a.f[i] = fx;
    a.f[(size_t)a.cap + i] = fy;
    a.f[2 * (size_t)a.cap + i] = fz;
}
template <int U4, int RCP>
__global__ void __launch_bounds__( 128 ) k_d4( const __grid_constant__ ArgsD a, const int4 *nb4, int rows4 )
{
    atom_d4<U4, RCP>( a, nb4, rows4, blockIdx.x * blockDim.x + threadIdx.x );
}
template <int U4, int RCP>
__global__ void __launch_bounds__( 128 )
    k_d4p( const __grid_constant__ ArgsD a, const int4 *nb4, int rows4, int *counters, int nsm, int n_tiles, const int *order )
{
    const int lane = threadIdx.x & 31;
    persistent_tiles( counters, nsm, n_tiles, order, [&]( int t ) { atom_d4<U4, RCP>( a, nb4, rows4, t * 32 + lane ); } );
}
__global__ void k_pack4( const int *__restrict__ nb, const int *__restrict__ cnt, int rows, int n, int4 *__restrict__ nb4, int rows4 )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const int c = cnt[i];
    for ( int k4 = 0; k4 < ( c + 3 ) / 4; k4++ )
    {
        int v[4];
        for ( int u = 0; u < 4; u++ )
            v[u] = 4 * k4 + u < c ? nb[TB( i, rows ) + (size_t)( 4 * k4 + u ) * 32] : i;
        nb4[( (size_t)( i >> 5 ) * rows4 + k4 ) * 32 + ( i & 31 )] = make_int4( v[0], v[1], v[2], v[3] );
    }
}

// ----------------------------------------------------------------------------- FP32 pieces
struct ArgsS
{
    const float4 *xf; // {x, y, z, type bits}
    cudaTextureObject_t tex4;
    const float2 *xyf;
    cudaTextureObject_t texzf;
    const int *nb;
    const int *cnt;
    int rows, n;
    float *f; // [3][cap]
    int cap;
    float lj1, lj2, cutsq;
};

// FETCH 0: LDG.128 float4; 1: TEX float4; 2: alternate LDG / TEX by k; 3: float2 LDG.64 + z TEX (4 B texel)
template <int U, int FETCH, int IDXH>
__device__ __forceinline__ void atom_s( const ArgsS &a, int i )
{
    if ( i >= a.n )
        return;
    const float4 xi = __ldg( a.xf + i );
    float fx = 0, fy = 0, fz = 0;
    const int c = a.cnt[i];
    const int *p = a.nb + TB( i, a.rows );
    const float lj1 = a.lj1, lj2 = a.lj2, cutsq = a.cutsq;
    const unsigned long long pf = policy_evict_first(), pl = policy_evict_last();
#pragma unroll( U )
    for ( int k = 0; k < c; k++ )
    {
        constexpr int H0 = IDXH < 3 ? IDXH : 0;
        const int j = IDXH >= 3 ? ld_idx_pol( p + k * 32, pf ) : ld_idx<H0>( p + k * 32 );
        float xj, yj, zj;
        if ( FETCH == 0 || ( FETCH == 2 && ( k & 1 ) == 0 ) )
        {
            const float4 t = IDXH == 4 ? ld_f4_pol( a.xf + j, pl ) : __ldg( a.xf + j );
            xj = t.x, yj = t.y, zj = t.z;
        }
        else if ( FETCH == 1 || FETCH == 2 )
        {
            const float4 t = tex1Dfetch<float4>( a.tex4, j );
            xj = t.x, yj = t.y, zj = t.z;
        }
        else
        {
            const float2 t = __ldg( a.xyf + j );
            xj = t.x, yj = t.y;
            zj = tex1Dfetch<float>( a.texzf, j );
        }
        const float dx = xi.x - xj, dy = xi.y - yj, dz = xi.z - zj;
        const float rsq = dx * dx + dy * dy + dz * dz;
        if ( rsq < cutsq )
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
    a.f[i] = fx;
    a.f[(size_t)a.cap + i] = fy;
    a.f[2 * (size_t)a.cap + i] = fz;
}
template <int U, int FETCH, int IDXH>
__global__ void __launch_bounds__( 128 ) k_s( const __grid_constant__ ArgsS a )
{
    atom_s<U, FETCH, IDXH>( a, blockIdx.x * blockDim.x + threadIdx.x );
}
template <int U, int BS>
__global__ void __launch_bounds__( BS ) k_sb( const __grid_constant__ ArgsS a )
{
    atom_s<U, 0, 0>( a, blockIdx.x * blockDim.x + threadIdx.x );
}

// bank-aware row order (round 1, experiments/force_variants.cu): reorder every row so that at
// step r lane q (= i & (NC-1)) reads a neighbour of class (r + q) & (NC-1), class = j & (NC-1)
template <int NC>
__global__ void __launch_bounds__( 128 ) k_reorder_rows( int *__restrict__ nb, const int *__restrict__ cnt, int rows, int n )
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
        const int j = nb[TB( i, rows ) + (size_t)k * 32];
        const int b = j & ( NC - 1 );
        if ( nc[b] < 128 / NC + 16 )
            cls[b][nc[b]++] = j;
    }
    const int q = i & ( NC - 1 );
    for ( int r = 0; r < c; r++ )
    {
        int b = ( r + q ) & ( NC - 1 );
        if ( used[b] >= nc[b] )
        {
            int best = -1, left = 0;
            for ( int t = 0; t < NC; t++ )
                if ( nc[t] - used[t] > left )
                {
                    left = nc[t] - used[t];
                    best = t;
                }
            b = best;
        }
        nb[TB( i, rows ) + (size_t)r * 32] = cls[b][used[b]++];
    }
}
template <int U, int FETCH, int IDXH>
__global__ void __launch_bounds__( 128 )
    k_sp( const __grid_constant__ ArgsS a, int *counters, int nsm, int n_tiles, const int *order )
{
    const int lane = threadIdx.x & 31;
    persistent_tiles( counters, nsm, n_tiles, order, [&]( int t ) { atom_s<U, FETCH, IDXH>( a, t * 32 + lane ); } );
}

// FP32, int4 index loads: IDXTEX 0 = LDG.128 from nb4, 1 = TEX int4 texel (index stream off the
// LSU pipe); TEXMASK bit u set = gather u of each group of 4 goes through TEX instead of LDG.128
template <int U4, int IDXTEX, int TEXMASK>
__device__ __forceinline__ void atom_s4( const ArgsS &a, const int4 *nb4, cudaTextureObject_t texnb, int rows4, int i )
{
    if ( i >= a.n )
        return;
    const float4 xi = __ldg( a.xf + i );
    float fx = 0, fy = 0, fz = 0;
    const int c = a.cnt[i];
    const int base = ( ( i >> 5 ) * rows4 ) * 32 + ( i & 31 );
    const float lj1 = a.lj1, lj2 = a.lj2, cutsq = a.cutsq;
    const int c4 = ( c + 3 ) >> 2;
#pragma unroll( U4 )
    for ( int k4 = 0; k4 < c4; k4++ )
    {
        const int4 q = IDXTEX ? tex1Dfetch<int4>( texnb, base + k4 * 32 ) : __ldg( nb4 + base + k4 * 32 );
        const int jj[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
        {
            const int j = jj[u]; // rows are padded with the atom itself: rsq = 0 -> rejected
            const float4 t = ( TEXMASK >> u ) & 1 ? tex1Dfetch<float4>( a.tex4, j ) : __ldg( a.xf + j );
            const float dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - t.z;
            const float rsq = dx * dx + dy * dy + dz * dz;
            if ( rsq < cutsq && j != i )
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
    a.f[i] = fx;
    a.f[(size_t)a.cap + i] = fy;
    a.f[2 * (size_t)a.cap + i] = fz;
}
template <int U4, int IDXTEX, int TEXMASK>
__global__ void __launch_bounds__( 128 )
    k_s4( const __grid_constant__ ArgsS a, const int4 *nb4, cudaTextureObject_t texnb, int rows4 )
{
    atom_s4<U4, IDXTEX, TEXMASK>( a, nb4, texnb, rows4, blockIdx.x * blockDim.x + threadIdx.x );
}

// FP64, index stream through TEX (int4 texels), xy LDG.128 + z TEX as in the product
template <int U4, int RCP, int IDXTEX>
__global__ void __launch_bounds__( 128 )
    k_d4t( const __grid_constant__ ArgsD a, const int4 *nb4, cudaTextureObject_t texnb, int rows4 )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= a.n )
        return;
    const XT xi = ld_xt( a.xt + i );
    double fx = 0, fy = 0, fz = 0;
    const int c = a.cnt[i];
    const int base = ( ( i >> 5 ) * rows4 ) * 32 + ( i & 31 );
    const double lj1 = a.lj1, lj2 = a.lj2, cutsq = a.cutsq;
    const int c4 = ( c + 3 ) >> 2;
#pragma unroll( U4 )
    for ( int k4 = 0; k4 < c4; k4++ )
    {
        const int4 q = IDXTEX ? tex1Dfetch<int4>( texnb, base + k4 * 32 ) : __ldg( nb4 + base + k4 * 32 );
        const int jj[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
        {
            const int j = jj[u];
            const double2 t = __ldg( a.xy + j );
            const int2 d = tex1Dfetch<int2>( a.texz, j );
            const double dx = xi.x - t.x, dy = xi.y - t.y, dz = xi.z - __hiloint2double( d.y, d.x );
            const double rsq = dx * dx + dy * dy + dz * dz;
            if ( lt_pos( rsq, cutsq ) && j != i )
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
    a.f[i] = fx;
    a.f[(size_t)a.cap + i] = fy;
    a.f[2 * (size_t)a.cap + i] = fz;
}

// Fast bank-aware row order on the int4-packed table, one warp per 32-atom tile, in place.
// Row of lane q: entries are dealt in ROUNDS; round m takes the m-th entry of every class
// (class = (j - q) & 7, i.e. relative to the lane) that still has one, in class order.  While
// all 8 classes last, slot r holds class (r + q) & 7: the 8 lanes of an LDG.128 group read 8
// different 16-byte positions of their 128-byte lines.  Position of the m-th entry of
// relative class c:  sum_b min(size_b, m)  +  #{b < c : size_b > m}  (byte-SIMD on packed sizes).
__global__ void __launch_bounds__( 128 )
    k_reorder_rounds( int4 *__restrict__ nb4, const int *__restrict__ cnt, int rows4, int n )
{
    extern __shared__ int sm[]; // [4 warps][rows4*4][32]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * 4 + w;
    const int i = tile * 32 + lane;
    if ( tile * 32 >= n )
        return;
    int *out = sm + (size_t)w * rows4 * 4 * 32;
    const int c = i < n ? cnt[i] : 0;
    int cmax = c;
    for ( int o = 16; o > 0; o >>= 1 )
        cmax = max( cmax, __shfl_xor_sync( 0xffffffffu, cmax, o ) );
    const int c4max = ( cmax + 3 ) >> 2;
    int4 *p = nb4 + ( (size_t)tile * rows4 ) * 32 + lane;
    const int q = i & 7;
    // pass 1: class sizes, packed bytes (relative classes 0-3 in lo, 4-7 in hi)
    unsigned lo = 0, hi = 0;
    for ( int k4 = 0; k4 < c4max; k4++ )
    {
        const int4 v = p[k4 * 32];
        const int jj[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
            if ( 4 * k4 + u < c )
            {
                const int cl = ( jj[u] - q ) & 7;
                if ( cl < 4 )
                    lo += 1u << ( 8 * cl );
                else
                    hi += 1u << ( 8 * ( cl - 4 ) );
            }
    }
    // pass 2: scatter into the staged column
    unsigned mlo = 0, mhi = 0; // running per-class counts
    for ( int k4 = 0; k4 < c4max; k4++ )
    {
        const int4 v = p[k4 * 32];
        const int jj[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for ( int u = 0; u < 4; u++ )
            if ( 4 * k4 + u < c )
            {
                const int cl = ( jj[u] - q ) & 7;
                const unsigned sh = 8 * ( cl & 3 );
                const unsigned m = ( ( cl < 4 ? mlo : mhi ) >> sh ) & 0xffu;
                const unsigned mm = m * 0x01010101u;
                int pos = __vsadu4( __vminu4( lo, mm ), 0 ) + __vsadu4( __vminu4( hi, mm ), 0 );
                const unsigned glo = __vcmpgtu4( lo, mm ), ghi = __vcmpgtu4( hi, mm );
                const unsigned long long g = ( (unsigned long long)ghi << 32 ) | glo;
                const unsigned long long mask = ( 1ull << ( 8 * cl ) ) - 1ull;
                pos += __popcll( g & mask ) >> 3;
                out[pos * 32 + lane] = jj[u];
                if ( cl < 4 )
                    mlo += 1u << sh;
                else
                    mhi += 1u << sh;
            }
    }
    // pad to a multiple of 4 with the atom itself
    for ( int k = c; k < ( ( c + 3 ) & ~3 ); k++ )
        out[k * 32 + lane] = i;
    __syncwarp();
    const int c4 = ( c + 3 ) >> 2;
    for ( int k4 = 0; k4 < c4; k4++ )
        p[k4 * 32] = make_int4( out[( 4 * k4 ) * 32 + lane], out[( 4 * k4 + 1 ) * 32 + lane],
                                out[( 4 * k4 + 2 ) * 32 + lane], out[( 4 * k4 + 3 ) * 32 + lane] );
}

// FP64 reference on arbitrary XT positions (plain LDG.256 kernel)
__global__ void __launch_bounds__( 128 )
    k_ref( const XT *__restrict__ xt, const int *__restrict__ nb, const int *__restrict__ cnt, int rows, int n,
           double *__restrict__ f, int cap, double lj1, double lj2, double cutsq, unsigned long long *inrange )
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n )
        return;
    const XT xi = xt[i];
    double fx = 0, fy = 0, fz = 0;
    const int c = cnt[i];
    int in = 0;
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
            in++;
        }
    }
    f[i] = fx;
    f[(size_t)cap + i] = fy;
    f[2 * (size_t)cap + i] = fz;
    if ( inrange )
        atomicAdd( inrange, (unsigned long long)in );
}

int main( int argc, char **argv )
{
    const int cells = argc > 1 ? atoi( argv[1] ) : 100;
    const int reps = argc > 2 ? atoi( argv[2] ) : 5;
    const std::string filter = argc > 3 ? argv[3] : "";
    MdSetup S;
    S.build( cells, 2.5, 0.3 );
    const int n = S.n, ntot = S.ntot, cap = S.cap;
    cudaDeviceProp prop;
    CK( cudaGetDeviceProperties( &prop, 0 ) );
    const int nsm = prop.multiProcessorCount;
    printf( "device %s, %d SMs\n", prop.name, nsm );

    // device data ----------------------------------------------------------------------------
    std::vector<XT> hxt( cap ), hxt32( cap );
    std::vector<double2> hxy( cap );
    std::vector<double> hz( cap );
    std::vector<float4> hxf( cap );
    std::vector<float2> hxyf( cap );
    std::vector<float> hzf( cap );
    for ( int i = 0; i < ntot; i++ )
    {
        const double *p = &S.x[3 * (size_t)i];
        hxt[i] = { p[0], p[1], p[2], 0 };
        hxy[i] = make_double2( p[0], p[1] );
        hz[i] = p[2];
        hxf[i] = make_float4( (float)p[0], (float)p[1], (float)p[2], 0.f );
        hxyf[i] = make_float2( (float)p[0], (float)p[1] );
        hzf[i] = (float)p[2];
        hxt32[i] = { (double)hxf[i].x, (double)hxf[i].y, (double)hxf[i].z, 0 };
    }
    XT *d_xt, *d_xt32;
    double2 *d_xy;
    double *d_z, *d_f, *d_ref;
    float4 *d_xf;
    float2 *d_xyf;
    float *d_zf, *d_fs;
    int *d_cs, *d_ca, *d_nb, *d_cnt, *d_counters, *d_order;
    const int rows = 128, n32 = ( n + 31 ) & ~31, n_tiles = n32 / 32;
    CK( cudaMalloc( &d_xt, (size_t)cap * sizeof( XT ) ) );
    CK( cudaMalloc( &d_xt32, (size_t)cap * sizeof( XT ) ) );
    CK( cudaMalloc( &d_xy, (size_t)cap * 24 ) ); // xy | z in one allocation (one L2 access-policy window)
    d_z = (double *)( d_xy + cap );
    CK( cudaMalloc( &d_xf, (size_t)cap * 16 ) );
    CK( cudaMalloc( &d_xyf, (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_zf, (size_t)cap * 4 ) );
    CK( cudaMalloc( &d_f, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_ref, 3 * (size_t)cap * 8 ) );
    CK( cudaMalloc( &d_fs, 3 * (size_t)cap * 4 ) );
    CK( cudaMalloc( &d_cs, ( S.cell_start.size() ) * 4 ) );
    CK( cudaMalloc( &d_ca, (size_t)ntot * 4 ) );
    CK( cudaMalloc( &d_nb, (size_t)rows * n32 * 4 ) );
    CK( cudaMalloc( &d_cnt, (size_t)cap * 4 ) );
    CK( cudaMalloc( &d_counters, 1024 * 4 ) );
    CK( cudaMalloc( &d_order, (size_t)n_tiles * 4 ) );
    CK( cudaMemcpy( d_xt, hxt.data(), (size_t)cap * sizeof( XT ), cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xt32, hxt32.data(), (size_t)cap * sizeof( XT ), cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xy, hxy.data(), (size_t)cap * 16, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_z, hz.data(), (size_t)cap * 8, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xf, hxf.data(), (size_t)cap * 16, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_xyf, hxyf.data(), (size_t)cap * 8, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_zf, hzf.data(), (size_t)cap * 4, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_cs, S.cell_start.data(), S.cell_start.size() * 4, cudaMemcpyHostToDevice ) );
    CK( cudaMemcpy( d_ca, S.cell_atoms.data(), (size_t)ntot * 4, cudaMemcpyHostToDevice ) );
    k_build_list<false><<<( n + 127 ) / 128, 128>>>( d_xt, n, d_cs, d_ca, S.nc, S.mn, S.rdx, S.rn * S.rn, d_nb, rows, d_cnt );
    CK( cudaDeviceSynchronize() );
    std::vector<int> hcnt( n );
    CK( cudaMemcpy( hcnt.data(), d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost ) );
    long long tot = 0;
    int mxc = 0;
    for ( int c : hcnt )
        tot += c, mxc = std::max( mxc, c );
    const double nn = (double)tot / n;

    // blob order of the tiles: 4x4x4-cell blocks, inside a block column by column
    {
        std::vector<int> order( n_tiles );
        std::vector<long long> key( n_tiles );
        for ( int t = 0; t < n_tiles; t++ )
        {
            const int c = S.acell[std::min( t * 32 + 16, n - 1 )];
            const int cz = c % S.nc, cy = ( c / S.nc ) % S.nc, cx = c / ( S.nc * S.nc );
            key[t] = ( ( ( ( (long long)( cx >> 2 ) * 64 + ( cy >> 2 ) ) * 64 + ( cz >> 2 ) ) * 4 + ( cx & 3 ) ) * 4 + ( cy & 3 ) ) * 1024 + cz;
        }
        std::iota( order.begin(), order.end(), 0 );
        std::stable_sort( order.begin(), order.end(), [&]( int p, int q ) { return key[p] < key[q]; } );
        CK( cudaMemcpy( d_order, order.data(), (size_t)n_tiles * 4, cudaMemcpyHostToDevice ) );
    }

    // textures
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
    cudaTextureObject_t tex4 = make_tex( d_xf, (size_t)cap * 16, cudaCreateChannelDesc<float4>() );
    cudaTextureObject_t texzf = make_tex( d_zf, (size_t)cap * 4, cudaCreateChannelDesc<float>() );

    const double lj1 = 48.0, lj2 = 24.0, cutsq = S.rc * S.rc;
    const int grid = ( n + 127 ) / 128;
    std::vector<double> ref( 3 * (size_t)cap ), ref32( 3 * (size_t)cap ), got( 3 * (size_t)cap );
    std::vector<float> gots( 3 * (size_t)cap );
    unsigned long long *d_in, h_in = 0;
    CK( cudaMalloc( &d_in, 8 ) );
    CK( cudaMemset( d_in, 0, 8 ) );
    k_ref<<<grid, 128>>>( d_xt, d_nb, d_cnt, rows, n, d_ref, cap, lj1, lj2, cutsq, d_in );
    CK( cudaMemcpy( ref.data(), d_ref, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
    CK( cudaMemcpy( &h_in, d_in, 8, cudaMemcpyDeviceToHost ) );
    k_ref<<<grid, 128>>>( d_xt32, d_nb, d_cnt, rows, n, d_ref, cap, lj1, lj2, cutsq, nullptr );
    CK( cudaMemcpy( ref32.data(), d_ref, 3 * (size_t)cap * 8, cudaMemcpyDeviceToHost ) );
    printf( "neighbours/atom %.2f (max %d), inside the force cutoff %.2f\n", nn, mxc, (double)h_in / n );
    const double bytes64 = n * ( 4.0 * nn + 32.0 ) + (double)ntot * 28.0;
    const double bytes32 = n * ( 4.0 * nn + 12.0 + 8.0 ) + (double)ntot * 16.0;

    Timer T;
    auto report = [&]( const char *name, float ms, double bytes, double err )
    {
        printf( "%-44s %8.4f ms  %7.1f GB/s alg (%.3f of 6547.5)  relerr %.2e\n", name, ms, bytes / ( ms * 1e-3 ) / 1e9,
                bytes / ( ms * 1e-3 ) / 1e9 / 6547.5, err );
        fflush( stdout );
    };
    auto want = [&]( const char *name ) { return filter.empty() || strstr( name, filter.c_str() ) != nullptr; };
    auto run_d = [&]( const char *name, auto launch )
    {
        if ( !want( name ) )
            return;
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
        if ( !want( name ) )
            return;
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

    ArgsD ad = { d_xt, d_xy, texz, d_nb, d_cnt, rows, n, d_f, cap, lj1, lj2, cutsq };
    ArgsS as = { d_xf, tex4, d_xyf, texzf, d_nb, d_cnt, rows, n, d_fs, cap, (float)lj1, (float)lj2, (float)cutsq };

#define RUN_D( NAME, U, RCP, ICMP, IDXH ) run_d( NAME, [&] { k_d<U, RCP, ICMP, IDXH><<<grid, 128>>>( ad ); } )
    RUN_D( "d0 product: rcp5 dsetp u6", 6, 5, false, 0 );
    RUN_D( "d1 rcp3 icmp u6", 6, 3, true, 0 );
    RUN_D( "d1 rcp3 dsetp u6", 6, 3, false, 0 );
    RUN_D( "d1 rcp3 icmp u4", 4, 3, true, 0 );
    RUN_D( "d1 rcp3 icmp u8", 8, 3, true, 0 );
    RUN_D( "d2 rcp2 icmp u6", 6, 2, true, 0 );
    RUN_D( "d1 rcp3 icmp u6 idx no_allocate", 6, 3, true, 1 );
    RUN_D( "d1 rcp3 icmp u6 idx .cs", 6, 3, true, 2 );

    auto persist_grid = [&]( auto kernel )
    {
        int occ = 0;
        CK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &occ, kernel, 128, 0 ) );
        return occ;
    };
#define RUN_DP( NAME, U, RCP, ICMP, IDXH, ORDER, OCCDIV )                                         \
    run_d( NAME, [&] {                                                                            \
        const int occ = std::max( 1, persist_grid( k_dp<U, RCP, ICMP, IDXH> ) / OCCDIV );         \
        cudaMemsetAsync( d_counters, 0, nsm * 4 );                                                \
        k_dp<U, RCP, ICMP, IDXH><<<nsm * occ, 128>>>( ad, d_counters, nsm, n_tiles, ORDER );      \
    } )
    RUN_DP( "dp persistent natural rcp3 icmp u6", 6, 3, true, 0, nullptr, 1 );
    RUN_DP( "dp persistent natural rcp3 icmp u6 noalloc", 6, 3, true, 1, nullptr, 1 );
    RUN_DP( "dp persistent blob    rcp3 icmp u6", 6, 3, true, 0, d_order, 1 );
    RUN_DP( "dp persistent blob    rcp3 icmp u6 noalloc", 6, 3, true, 1, d_order, 1 );
    RUN_DP( "dp persistent blob    rcp3 icmp u4 noalloc", 4, 3, true, 1, d_order, 1 );
    RUN_DP( "dp persistent blob    rcp3 icmp u8 noalloc", 8, 3, true, 1, d_order, 1 );
    RUN_DP( "dp persistent blob    rcp5 dsetp u6", 6, 5, false, 0, d_order, 1 );
    RUN_DP( "dp persistent blob    rcp3 icmp u6 half-occ", 6, 3, true, 1, d_order, 2 );
    RUN_DP( "dp persistent natural rcp2 icmp u6 noalloc", 6, 2, true, 1, nullptr, 1 );

    {
        const int rows4 = rows / 4;
        int4 *d_nb4;
        CK( cudaMalloc( &d_nb4, (size_t)rows4 * n32 * 16 ) );
        k_pack4<<<grid, 128>>>( d_nb, d_cnt, rows, n, d_nb4, rows4 );
        CK( cudaDeviceSynchronize() );
        run_d( "d4 int4 index loads rcp3 u1x4", [&] { k_d4<1, 3><<<grid, 128>>>( ad, d_nb4, rows4 ); } );
        run_d( "d4 int4 index loads rcp3 u2x4", [&] { k_d4<2, 3><<<grid, 128>>>( ad, d_nb4, rows4 ); } );
        run_d( "d4p persistent blob int4 rcp3 u2x4", [&] {
            const int occ = persist_grid( k_d4p<2, 3> );
            cudaMemsetAsync( d_counters, 0, nsm * 4 );
            k_d4p<2, 3><<<nsm * occ, 128>>>( ad, d_nb4, rows4, d_counters, nsm, n_tiles, d_order );
        } );
        run_d( "d4p persistent natural int4 rcp3 u2x4", [&] {
            const int occ = persist_grid( k_d4p<2, 3> );
            cudaMemsetAsync( d_counters, 0, nsm * 4 );
            k_d4p<2, 3><<<nsm * occ, 128>>>( ad, d_nb4, rows4, d_counters, nsm, n_tiles, nullptr );
        } );
        CK( cudaFree( d_nb4 ) );
    }

#define RUN_S( NAME, U, FETCH, IDXH ) run_s( NAME, [&] { k_s<U, FETCH, IDXH><<<grid, 128>>>( as ); } )
    RUN_S( "s0 fp32 float4 LDG.128 u4", 4, 0, 0 );
    RUN_S( "s0 fp32 float4 LDG.128 u8", 8, 0, 0 );
    RUN_S( "s0 fp32 float4 LDG.128 u12", 12, 0, 0 );
    RUN_S( "s1 fp32 float4 TEX u8", 8, 1, 0 );
    RUN_S( "s2 fp32 alternate LDG/TEX u8", 8, 2, 0 );
    RUN_S( "s3 fp32 xy LDG.64 + z TEX u8", 8, 3, 0 );
    RUN_S( "s3 fp32 xy LDG.64 + z TEX u4", 4, 3, 0 );
#define RUN_SP( NAME, U, FETCH, IDXH, ORDER )                                                     \
    run_s( NAME, [&] {                                                                            \
        const int occ = persist_grid( k_sp<U, FETCH, IDXH> );                                     \
        cudaMemsetAsync( d_counters, 0, nsm * 4 );                                                \
        k_sp<U, FETCH, IDXH><<<nsm * occ, 128>>>( as, d_counters, nsm, n_tiles, ORDER );          \
    } )
    RUN_SP( "sp persistent natural float4 LDG.128 u8", 8, 0, 0, nullptr );
    RUN_SP( "sp persistent blob    float4 LDG.128 u8", 8, 0, 0, d_order );
    RUN_SP( "sp persistent blob    float4 LDG.128 u8 noalloc", 8, 0, 1, d_order );
    RUN_SP( "sp persistent blob    alternate u8 noalloc", 8, 2, 1, d_order );
    RUN_SP( "sp persistent blob    xy LDG.64 + z TEX u8 noalloc", 8, 3, 1, d_order );
    RUN_SP( "sp persistent blob    float4 LDG.128 u4 noalloc", 4, 0, 1, d_order );

    // ---- batch 2: L2 residency of the positions, fine unroll / block size, bank-aware rows
    RUN_D( "e1 d1 + idx L2 evict_first", 6, 3, true, 3 );
    RUN_D( "e2 d1 + idx evict_first + xy evict_last", 6, 3, true, 4 );
    RUN_S( "e1 s0 u4 + idx L2 evict_first", 4, 0, 3 );
    RUN_S( "e2 s0 u4 + idx evict_first + pos evict_last", 4, 0, 4 );
    RUN_S( "s0 fp32 float4 LDG.128 u2", 2, 0, 0 );
    RUN_S( "s0 fp32 float4 LDG.128 u3", 3, 0, 0 );
    RUN_S( "s0 fp32 float4 LDG.128 u5", 5, 0, 0 );
    RUN_S( "s0 fp32 float4 LDG.128 u6", 6, 0, 0 );
    run_s( "sb fp32 u4 bs64", [&] { k_sb<4, 64><<<( n + 63 ) / 64, 64>>>( as ); } );
    run_s( "sb fp32 u4 bs256", [&] { k_sb<4, 256><<<( n + 255 ) / 256, 256>>>( as ); } );
    run_s( "sb fp32 u4 bs512", [&] { k_sb<4, 512><<<( n + 511 ) / 512, 512>>>( as ); } );
    {
        printf( "persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB, L2 %d MB\n", prop.persistingL2CacheMaxSize >> 20,
                prop.accessPolicyMaxWindowSize >> 20, prop.l2CacheSize >> 20 );
        CK( cudaDeviceSetLimit( cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize ) );
        auto window = [&]( void *ptr, size_t bytes, float ratio )
        {
            cudaStreamAttrValue av = {};
            av.accessPolicyWindow.base_ptr = ptr;
            av.accessPolicyWindow.num_bytes = std::min( bytes, (size_t)prop.accessPolicyMaxWindowSize );
            av.accessPolicyWindow.hitRatio = ratio;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK( cudaStreamSetAttribute( 0, cudaStreamAttributeAccessPolicyWindow, &av ) );
        };
        for ( float ratio : { 1.0f, 0.6f } )
        {
            char name[96];
            window( d_xy, (size_t)cap * 24, ratio );
            snprintf( name, sizeof name, "e3 d1 + persisting window xy|z ratio %.1f", ratio );
            RUN_D( name, 6, 3, true, 0 );
            snprintf( name, sizeof name, "e3 d1 + window %.1f + idx evict_first", ratio );
            RUN_D( name, 6, 3, true, 3 );
            window( d_xf, (size_t)cap * 16, ratio );
            snprintf( name, sizeof name, "e3 s0 u4 + persisting window ratio %.1f", ratio );
            RUN_S( name, 4, 0, 0 );
            snprintf( name, sizeof name, "e3 s0 u4 + window %.1f + idx evict_first", ratio );
            RUN_S( name, 4, 0, 3 );
        }
        cudaStreamAttrValue av = {};
        av.accessPolicyWindow.num_bytes = 0;
        CK( cudaStreamSetAttribute( 0, cudaStreamAttributeAccessPolicyWindow, &av ) );
        CK( cudaCtxResetPersistingL2Cache() );
    }
    const int rows4 = rows / 4;
    int4 *d_nb4;
    CK( cudaMalloc( &d_nb4, (size_t)rows4 * n32 * 16 ) );
    cudaTextureObject_t texnb = make_tex( d_nb4, (size_t)rows4 * n32 * 16, cudaCreateChannelDesc<int4>() );
    auto batch3 = [&]( const char *tag )
    {
        k_pack4<<<grid, 128>>>( d_nb, d_cnt, rows, n, d_nb4, rows4 );
        CK( cudaDeviceSynchronize() );
        char name[96];
#define RUN_S4( LABEL, U4, IT, TM )                                                               \
    snprintf( name, sizeof name, "g %s %s", LABEL, tag );                                         \
    run_s( name, [&] { k_s4<U4, IT, TM><<<grid, 128>>>( as, d_nb4, texnb, rows4 ); } )
        RUN_S4( "s4 idx LDG.128, gathers LDG u1x4", 1, 0, 0 );
        RUN_S4( "s4 idx LDG.128, gathers LDG u2x4", 2, 0, 0 );
        RUN_S4( "s4 idx TEX, gathers LDG u1x4", 1, 1, 0 );
        RUN_S4( "s4 idx TEX, gathers LDG u2x4", 2, 1, 0 );
        RUN_S4( "s4 idx LDG.128, gathers 3 LDG + 1 TEX u1x4", 1, 0, 8 );
        RUN_S4( "s4 idx LDG.128, gathers 3 LDG + 1 TEX u2x4", 2, 0, 8 );
        RUN_S4( "s4 idx TEX, gathers 3 LDG + 1 TEX u1x4", 1, 1, 8 );
        RUN_S4( "s4 idx TEX, gathers 3 LDG + 1 TEX u2x4", 2, 1, 8 );
        RUN_S4( "s4 idx LDG.128, gathers 2 LDG + 2 TEX u1x4", 1, 0, 10 );
        snprintf( name, sizeof name, "g d4t idx TEX, xy LDG.128 + z TEX u1x4 %s", tag );
        run_d( name, [&] { k_d4t<1, 3, 1><<<grid, 128>>>( ad, d_nb4, texnb, rows4 ); } );
        snprintf( name, sizeof name, "g d4t idx TEX, xy LDG.128 + z TEX u2x4 %s", tag );
        run_d( name, [&] { k_d4t<2, 3, 1><<<grid, 128>>>( ad, d_nb4, texnb, rows4 ); } );
        snprintf( name, sizeof name, "g d4t idx LDG.128, xy LDG.128 + z TEX u2x4 %s", tag );
        run_d( name, [&] { k_d4t<2, 3, 0><<<grid, 128>>>( ad, d_nb4, texnb, rows4 ); } );
        snprintf( name, sizeof name, "g d4t idx LDG.128, xy LDG.128 + z TEX u3x4 %s", tag );
        run_d( name, [&] { k_d4t<3, 3, 0><<<grid, 128>>>( ad, d_nb4, texnb, rows4 ); } );
        snprintf( name, sizeof name, "g d4t idx LDG.128, rcp2 u2x4 %s", tag );
        run_d( name, [&] { k_d4t<2, 2, 0><<<grid, 128>>>( ad, d_nb4, texnb, rows4 ); } );
        RUN_S4( "s4 idx TEX, gathers LDG u3x4", 3, 1, 0 );
        RUN_S4( "s4 idx TEX, gathers LDG u4x4", 4, 1, 0 );
    };
    batch3( "(index-ordered rows)" );
    // bank-aware rows last: they permute the table in place
    {
        const float ms = T.time( [&] { k_reorder_rows<8><<<grid, 128>>>( d_nb, d_cnt, rows, n ); }, 3, 1 );
        printf( "g k_reorder_rows<8>: %.3f ms per pass\n", ms );
    }
    batch3( "(rows 8-class)" );
    {
        // the fast round-based order, applied to a fresh index-ordered table
        k_build_list<false><<<( n + 127 ) / 128, 128>>>( d_xt, n, d_cs, d_ca, S.nc, S.mn, S.rdx, S.rn * S.rn, d_nb, rows, d_cnt );
        k_pack4<<<grid, 128>>>( d_nb, d_cnt, rows, n, d_nb4, rows4 );
        CK( cudaDeviceSynchronize() );
        const size_t smb = (size_t)4 * rows4 * 4 * 32 * 4;
        CK( cudaFuncSetAttribute( k_reorder_rounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb ) );
        const float ms = T.time( [&] { k_reorder_rounds<<<( n_tiles + 3 ) / 4, 128, smb>>>( d_nb4, d_cnt, rows4, n ); }, 3, 1 );
        printf( "g k_reorder_rounds: %.3f ms per pass (rows4 %d, %zu B smem per CTA)\n", ms, rows4, smb );
        char name[96];
        const char *tag = "(rows: rounds)";
        RUN_S4( "s4 idx TEX, gathers LDG u3x4", 3, 1, 0 );
        RUN_S4( "s4 idx TEX, gathers LDG u2x4", 2, 1, 0 );
        snprintf( name, sizeof name, "g d4t idx LDG.128, xy LDG.128 + z TEX u2x4 %s", tag );
        run_d( name, [&] { k_d4t<2, 3, 0><<<grid, 128>>>( ad, d_nb4, texnb, rows4 ); } );
    }
    if ( filter == "g " )
        return 0;
    k_reorder_rows<16><<<grid, 128>>>( d_nb, d_cnt, rows, n );
    CK( cudaDeviceSynchronize() );
    batch3( "(rows 16-class)" );
    k_reorder_rows<32><<<grid, 128>>>( d_nb, d_cnt, rows, n );
    CK( cudaDeviceSynchronize() );
    batch3( "(rows 32-class)" );
    RUN_S( "f8 s0 u4, rows 8-class", 4, 0, 0 );
    RUN_S( "f8 s0 u8, rows 8-class", 8, 0, 0 );
    RUN_D( "f8 d1 u6, rows 8-class", 6, 3, true, 0 );
    return 0;
}
