// experiments/ldg_patterns.cu — how the L1 LSU pipe splits one warp-wide gather into
// wavefronts, for 32-, 64-, 128- and 256-bit loads (sm_100a).  Every lane of every warp
// repeatedly loads from a small L1-resident array with a fixed lane->address pattern; the
// time per request (in SM cycles at the measured clock) is the wavefront cost of the
// pattern.  NOT part of the product or the tests.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o experiments/ldg_patterns experiments/ldg_patterns.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
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

// byte offset of lane `l` for pattern `p`; `w` = access width in bytes.  All offsets stay
// inside a 4 KiB window (32 lines) so everything hits L1 after the first touch.
__host__ __device__ inline int pattern_offset( int p, int l, int w )
{
    switch ( p )
    {
    case 0: return 0;                                   // broadcast: one element
    case 1: return l * w;                               // contiguous
    case 2: return l * 128;                             // 32 distinct lines
    case 3: return ( l >> 1 ) * 128 + ( l & 1 ) * w;    // adjacent pairs share a line (16 lines)
    case 4: return ( l >> 2 ) * 128 + ( l & 3 ) * w;    // adjacent quads share a line (8 lines)
    case 5: return ( l >> 3 ) * 128 + ( ( ( l & 7 ) * w ) & 127 );    // adjacent octets share a line (4 lines)
    case 6: return ( l & 15 ) * 128 + ( l >> 4 ) * w;   // lanes l, l+16 share a line (16 lines)
    case 7: return ( l & 7 ) * 128 + ( l >> 3 ) * w;    // lanes l, l+8, l+16, l+24 share (8 lines)
    case 8: return ( l & 3 ) * 128 + ( ( l >> 2 ) & 3 ) * ( w > 32 ? 32 : w ) + 0 * l; // lanes l, l+4.. share (4 lines)
    case 9: return ( ( l * 7 ) & 15 ) * 128 + ( ( l * 5 ) & 3 ) * w; // scrambled: 16 lines, 2 lanes each
    default: return 0;
    }
}

template <int W>
__device__ __forceinline__ unsigned long long load_w( const char *p )
{
    if ( W == 4 )
    {
        unsigned v;
        asm volatile( "ld.global.ca.u32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
        return v;
    }
    if ( W == 8 )
    {
        unsigned long long v;
        asm volatile( "ld.global.ca.u64 %0, [%1];" : "=l"( v ) : "l"( p ) : "memory" );
        return v;
    }
    if ( W == 16 )
    {
        unsigned long long a, b;
        asm volatile( "ld.global.ca.v2.u64 {%0,%1}, [%2];" : "=l"( a ), "=l"( b ) : "l"( p ) : "memory" );
        return a ^ b;
    }
    unsigned long long a, b, c, d;
    asm volatile( "ld.global.ca.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"( a ), "=l"( b ), "=l"( c ), "=l"( d ) : "l"( p ) : "memory" );
    return a ^ b ^ c ^ d;
}

template <int W>
__global__ void __launch_bounds__( 256 ) k_pattern( const char *base, int p, int iters, unsigned long long *sink )
{
    const int lane = threadIdx.x & 31;
    // 16 windows of 4 KiB (64 KiB, L1 resident); every iteration of every warp reads another
    // window so that no two loads in flight are the same instruction + address
    const int w0 = blockIdx.x * 8 + ( threadIdx.x >> 5 );
    const char *q = base + pattern_offset( p, lane, W );
    unsigned long long acc = 0;
#pragma unroll 8
    for ( int i = 0; i < iters; i++ )
        acc += load_w<W>( q + ( ( w0 + i ) & 15 ) * 4096 );
    if ( acc == 0x1234567ull )
        sink[0] = acc;
}

int main()
{
    char *buf;
    unsigned long long *sink;
    CK( cudaMalloc( &buf, 64 * 4096 + 4096 ) );
    CK( cudaMemset( buf, 1, 64 * 4096 + 4096 ) );
    CK( cudaMalloc( &sink, 8 ) );
    cudaDeviceProp prop;
    CK( cudaGetDeviceProperties( &prop, 0 ) );
    const int nsm = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute( &clk_khz, cudaDevAttrClockRate, 0 );
    const int iters = 4096, blocks = nsm * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate( &e0 );
    cudaEventCreate( &e1 );
    const char *names[10] = { "broadcast", "contiguous", "32 lines", "pairs share", "quads share", "octets share",
                              "l,l+16 share", "l,l+8,.. share", "l,l+4,.. share", "scrambled 16x2" };
    printf( "SMs %d, clock %.0f MHz (attr); cycles per warp request per SM (all warps of an SM share one LSU pipe)\n", nsm, clk_khz / 1e3 );
    printf( "%-16s %10s %10s %10s %10s\n", "pattern", "LDG.32", "LDG.64", "LDG.128", "LDG.256" );
    for ( int p = 0; p < 10; p++ )
    {
        printf( "%-16s", names[p] );
        for ( int w = 4; w <= 32; w *= 2 )
        {
            auto launch = [&]
            {
                if ( w == 4 )
                    k_pattern<4><<<blocks, 256>>>( buf, p, iters, sink );
                else if ( w == 8 )
                    k_pattern<8><<<blocks, 256>>>( buf, p, iters, sink );
                else if ( w == 16 )
                    k_pattern<16><<<blocks, 256>>>( buf, p, iters, sink );
                else
                    k_pattern<32><<<blocks, 256>>>( buf, p, iters, sink );
            };
            launch();
            CK( cudaDeviceSynchronize() );
            cudaEventRecord( e0 );
            launch();
            cudaEventRecord( e1 );
            CK( cudaDeviceSynchronize() );
            float ms;
            cudaEventElapsedTime( &ms, e0, e1 );
            // requests per SM = 4 CTAs x 8 warps x iters
            const double req_per_sm = 4.0 * 8.0 * iters;
            const double cyc = ms * 1e-3 * ( clk_khz * 1e3 ) / req_per_sm;
            printf( " %10.2f", cyc );
        }
        printf( "\n" );
    }
    return 0;
}
