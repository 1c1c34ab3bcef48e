// binary_dump.h — per-rank binary state dumps and the regression check against them:
// `--dumpbinary N PATH` and `--correctness N PATH FILE` (reference
// cabanamd_impl.h:434-642, option parsing inputCL.cpp:225-241).
//
// File PATH/output.<step, 10 digits>.<rank, 3 digits>, native endianness:
//   int32 n | int32 id[n] | int32 type[n] | float64 q[n] | float64 x[n][3] | v[n][3] | f[n][3]
// The check matches atoms by id and reports, per checked step, the l2 norm and the
// largest component of the position, velocity and force differences, reduced over ranks,
// as one line `step |dr| max|dr| |dv| max|dv| |df| max|df|` (%d %g ...) in FILE.
//
// In the reference both switches are parsed but never copied into the InputFile object
// that the step loop consults, and the checker indexes its (3,n) reference views as
// (i,0..2); the row-major [n][3] layout its dump writes is what is implemented here, with
// a hash map instead of the quadratic id search.  Host-only code: the CPU tests drive it
// through `cbmd_io_tool`.
#ifndef CBMD_HOST_BINARY_DUMP_H
#define CBMD_HOST_BINARY_DUMP_H

#include <cmath>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include "types.h"

namespace BinaryDump
{

struct State
{
    T_INT n = 0;
    std::vector<T_INT> id, type;
    std::vector<T_FLOAT> q;
    std::vector<T_X_FLOAT> x;
    std::vector<T_V_FLOAT> v;
    std::vector<T_F_FLOAT> f;
};

inline std::string file_name( const std::string &path, int step, int rank )
{
    char tail[64];
    std::snprintf( tail, sizeof tail, "/output.%010d.%03d", step, rank );
    return path + tail;
}

// first n rows of the host arrays -> file; false when the file cannot be written
inline bool write( const std::string &file, T_INT n, const T_INT *id, const T_INT *type, const T_FLOAT *q,
                   const T_X_FLOAT *x, const T_V_FLOAT *v, const T_F_FLOAT *f )
{
    FILE *fp = std::fopen( file.c_str(), "wb" );
    if ( !fp )
        return false;
    const size_t m = (size_t)n;
    bool ok = std::fwrite( &n, sizeof n, 1, fp ) == 1;
    ok = ok && std::fwrite( id, sizeof *id, m, fp ) == m;
    ok = ok && std::fwrite( type, sizeof *type, m, fp ) == m;
    ok = ok && std::fwrite( q, sizeof *q, m, fp ) == m;
    ok = ok && std::fwrite( x, sizeof *x, 3 * m, fp ) == 3 * m;
    ok = ok && std::fwrite( v, sizeof *v, 3 * m, fp ) == 3 * m;
    ok = ok && std::fwrite( f, sizeof *f, 3 * m, fp ) == 3 * m;
    return std::fclose( fp ) == 0 && ok;
}

enum ReadStatus
{
    READ_OK,
    READ_CANNOT_OPEN,
    READ_COUNT_MISMATCH, // the file's n differs from `expect_n` (>= 0)
    READ_SHORT
};

inline ReadStatus read( const std::string &file, T_INT expect_n, State &s )
{
    FILE *fp = std::fopen( file.c_str(), "rb" );
    if ( !fp )
        return READ_CANNOT_OPEN;
    ReadStatus st = READ_OK;
    if ( std::fread( &s.n, sizeof s.n, 1, fp ) != 1 || s.n < 0 )
        st = READ_SHORT;
    else if ( expect_n >= 0 && s.n != expect_n )
        st = READ_COUNT_MISMATCH;
    else
    {
        const size_t m = (size_t)s.n;
        s.id.resize( m ), s.type.resize( m ), s.q.resize( m );
        s.x.resize( 3 * m ), s.v.resize( 3 * m ), s.f.resize( 3 * m );
        bool ok = std::fread( s.id.data(), sizeof( T_INT ), m, fp ) == m;
        ok = ok && std::fread( s.type.data(), sizeof( T_INT ), m, fp ) == m;
        ok = ok && std::fread( s.q.data(), sizeof( T_FLOAT ), m, fp ) == m;
        ok = ok && std::fread( s.x.data(), sizeof( T_X_FLOAT ), 3 * m, fp ) == 3 * m;
        ok = ok && std::fread( s.v.data(), sizeof( T_V_FLOAT ), 3 * m, fp ) == 3 * m;
        ok = ok && std::fread( s.f.data(), sizeof( T_F_FLOAT ), 3 * m, fp ) == 3 * m;
        if ( !ok )
            st = READ_SHORT;
    }
    std::fclose( fp );
    return st;
}

// this rank's contribution to the report: sums of squared differences and largest
// absolute components (cabanamd_impl.h:545-590)
struct Deltas
{
    T_FLOAT sumsq[3] = { 0, 0, 0 }; // r, v, f
    T_FLOAT maxabs[3] = { 0, 0, 0 };
    T_INT unmatched_id = -1; // first reference id with no current atom, or -1
};

inline Deltas compare( T_INT n, const T_INT *id, const T_X_FLOAT *x, const T_V_FLOAT *v, const T_F_FLOAT *f,
                       const State &ref )
{
    Deltas d;
    std::unordered_map<T_INT, T_INT> row_of_id;
    row_of_id.reserve( (size_t)n * 2 );
    for ( T_INT i = n - 1; i >= 0; i-- ) // the lowest row wins, like the reference's linear search
        row_of_id[id[i]] = i;
    const T_FLOAT *cur[3] = { x, v, f };
    const T_FLOAT *old[3] = { ref.x.data(), ref.v.data(), ref.f.data() };
    for ( T_INT i = 0; i < ref.n; i++ )
    {
        T_INT ii = i < n && id[i] == ref.id[i] ? i : -1;
        if ( ii < 0 )
        {
            const auto it = row_of_id.find( ref.id[i] );
            if ( it == row_of_id.end() )
            {
                if ( d.unmatched_id < 0 )
                    d.unmatched_id = ref.id[i];
                continue;
            }
            ii = it->second;
        }
        for ( int a = 0; a < 3; a++ )
            for ( int c = 0; c < 3; c++ )
            {
                const T_FLOAT del = cur[a][3 * (size_t)ii + c] - old[a][3 * (size_t)i + c];
                d.sumsq[a] += del * del;
                d.maxabs[a] = std::fmax( std::fabs( del ), d.maxabs[a] );
            }
    }
    return d;
}

// one report line; the header goes in front of step 0, which also truncates the file
inline bool append_report( const std::string &file, int step, const T_FLOAT sumsq[3], const T_FLOAT maxabs[3] )
{
    FILE *fp = std::fopen( file.c_str(), step == 0 ? "w" : "a" );
    if ( !fp )
        return false;
    if ( step == 0 )
        std::fprintf( fp, "# timestep deltarnorm maxdelr deltavnorm maxdelv deltafnorm maxdelf\n" );
    std::fprintf( fp, "%d %g %g %g %g %g %g\n", step, std::sqrt( sumsq[0] ), maxabs[0], std::sqrt( sumsq[1] ),
                  maxabs[1], std::sqrt( sumsq[2] ), maxabs[2] );
    return std::fclose( fp ) == 0;
}

} // namespace BinaryDump

#endif
