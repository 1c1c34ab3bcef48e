// experiments/md_setup.h — host-side construction of a liquid-like LJ configuration with the
// product's data layout (cell-sorted owned atoms, 6-phase periodic ghosts, cell lists, tiled
// Verlet table).  Shared by the round-2 A/B harnesses; NOT part of the product or the tests.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
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

// tiled Verlet table of the product: neighbour k of atom i at nb[((i>>5)*rows + k)*32 + (i&31)]
#define TB( i, rows ) ( ( (size_t)( ( i ) >> 5 ) * (size_t)( rows ) ) * 32 + (size_t)( ( i ) & 31 ) )

struct alignas( 32 ) XT
{
    double x, y, z;
    long long t;
};

struct MdSetup
{
    int cells = 0, n = 0, ntot = 0, cap = 0;
    double L = 0, rn = 2.8, rc = 2.5;
    int nc = 0; // cells per dimension of the Verlet grid (incl. one ghost layer each side)
    double mn = 0, rdx = 0;
    std::vector<double> x;       // [ntot][3], owned atoms cell-sorted, ghosts appended
    std::vector<int> cell_start; // [nc^3+1]
    std::vector<int> cell_atoms; // [ntot]
    std::vector<int> acell;      // [ntot]

    int cell_of( const double *p ) const
    {
        int c[3];
        for ( int d = 0; d < 3; d++ )
        {
            int q = (int)std::floor( ( p[d] - mn ) * rdx );
            c[d] = std::min( std::max( q, 0 ), nc - 1 );
        }
        return ( c[0] * nc + c[1] ) * nc + c[2];
    }

    void build( int cells_, double rc_, double skin, double jitter = 0.22 )
    {
        cells = cells_;
        rc = rc_;
        rn = rc_ + skin;
        const double a = std::cbrt( 4.0 / 0.8442 );
        L = a * cells;
        n = 4 * cells * cells * cells;
        x.assign( 3 * (size_t)n, 0.0 );
        {
            const double basis[4][3] = { { 0, 0, 0 }, { .5, .5, 0 }, { .5, 0, .5 }, { 0, .5, .5 } };
            std::mt19937_64 rng( 12345 );
            std::uniform_real_distribution<double> u( -jitter, jitter );
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
        const int nbin = (int)( L / rn );
        const double dbin = L / nbin, eps = dbin / 1000;
        mn = -dbin - eps;
        const double mx = L + dbin + eps;
        nc = (int)std::floor( ( mx - mn ) / dbin );
        rdx = 1.0 / ( ( mx - mn ) / nc );
        {
            std::vector<int> cell( n ), order( n );
            for ( int i = 0; i < n; i++ )
                cell[i] = cell_of( &x[3 * (size_t)i] );
            std::iota( order.begin(), order.end(), 0 );
            std::stable_sort( order.begin(), order.end(), [&]( int p, int q ) { return cell[p] < cell[q]; } );
            std::vector<double> y( x.size() );
            for ( int i = 0; i < n; i++ )
                for ( int d = 0; d < 3; d++ )
                    y[3 * (size_t)i + d] = x[3 * (size_t)order[i] + d];
            x.swap( y );
        }
        size_t last_recv = 0;
        for ( int ph = 0; ph < 6; ph++ )
        {
            const int d = ph / 2;
            const size_t cur = x.size() / 3;
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
        ntot = (int)( x.size() / 3 );
        cap = ( ntot + 127 ) & ~127;
        const int ncells = nc * nc * nc;
        cell_start.assign( ncells + 1, 0 );
        cell_atoms.assign( ntot, 0 );
        acell.assign( ntot, 0 );
        for ( int i = 0; i < ntot; i++ )
        {
            acell[i] = cell_of( &x[3 * (size_t)i] );
            cell_start[acell[i] + 1]++;
        }
        for ( int c = 0; c < ncells; c++ )
            cell_start[c + 1] += cell_start[c];
        std::vector<int> cur( cell_start.begin(), cell_start.end() - 1 );
        for ( int i = 0; i < ntot; i++ )
            cell_atoms[cur[acell[i]]++] = i;
        printf( "cells %d  atoms %d  L %.3f  ghosts %d (%.1f%%)  grid %d^3\n", cells, n, L, ntot - n,
                100.0 * ( ntot - n ) / n, nc );
    }
};

// simple list builder: thread per atom over the 27-cell stencil, rows in ascending (cell, index)
template <bool HALF>
__global__ void k_build_list( const XT *__restrict__ xt, int n, const int *__restrict__ cell_start,
                              const int *__restrict__ cell_atoms, int nc, double mn, double rdx, double rsq,
                              int *__restrict__ nb, int rows, int *__restrict__ cnt )
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
            const int s0 = cell_start[row + max( c[2] - 1, 0 )], s1 = cell_start[row + min( c[2] + 1, nc - 1 ) + 1];
            for ( int s = s0; s < s1; s++ )
            {
                const int j = cell_atoms[s];
                const XT xj = xt[j];
                const double dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                bool ok = j != i && dx * dx + dy * dy + dz * dz <= rsq;
                if ( HALF )
                    ok = ok && ( xj.x > xi.x || ( xj.x == xi.x && ( xj.y > xi.y || ( xj.y == xi.y && xj.z > xi.z ) ) ) );
                if ( ok )
                {
                    if ( count < rows )
                        nb[TB( i, rows ) + count * 32] = j;
                    count++;
                }
            }
        }
    cnt[i] = count;
}

struct Timer
{
    cudaEvent_t e0, e1;
    Timer()
    {
        cudaEventCreate( &e0 );
        cudaEventCreate( &e1 );
    }
    template <class F>
    float time( F launch, int reps, int warm = 2 )
    {
        for ( int r = 0; r < warm; r++ )
            launch();
        CK( cudaDeviceSynchronize() );
        cudaEventRecord( e0 );
        for ( int r = 0; r < reps; r++ )
            launch();
        cudaEventRecord( e1 );
        CK( cudaDeviceSynchronize() );
        float ms;
        cudaEventElapsedTime( &ms, e0, e1 );
        return ms / reps;
    }
};
