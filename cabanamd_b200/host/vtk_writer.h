// vtk_writer.h — `dump ID all vtk N file_*%.vtu` particle dumps (reference
// src/vtk_writer.h:205-294, call site cabanamd_impl.h:388-390).
//
// File contents are the reference's byte for byte: an ASCII VTK UnstructuredGrid piece per
// rank with point data Velocity (3 x Float64), Id, Type (Int32), the Points array and
// three empty cell arrays, every number printed with printf's %g / %d followed by one
// blank; `*` in the pattern becomes the 4-digit zero-padded step and `%` becomes `_rank`.
// The .pvtu index is written by rank 1, as in the reference (:264-265) — a one-rank run
// has no index file.
//
// What is different is WHERE the work happens.  The reference formats ~8 numbers per atom
// with fprintf inside the step loop.  Here the step loop only takes a snapshot of the
// owned atoms (one device->host copy through the C ABI) and hands it to a background
// writer: the snapshot is cut into chunks that a small pool of threads format
// concurrently (std::to_chars, which is defined to produce printf's %g text), and the
// pieces are written in order.  The MD loop continues while a dump is being written; it
// only waits when a second dump arrives before the previous one has been taken over.
#ifndef CBMD_HOST_VTK_WRITER_H
#define CBMD_HOST_VTK_WRITER_H

#include <algorithm>
#include <charconv>
#include <condition_variable>
#include <cstdio>
#include <iomanip>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "output.h"
#include "types.h"

namespace VTKWriter
{

// step number padded like the reference's set_width (vtk_writer.h:23-28)
inline std::string step_field( int step, unsigned width = 4 )
{
    std::ostringstream s;
    s << std::setw( width ) << std::setfill( '0' ) << step;
    return s.str();
}

// resolves `*` and `%`; returns false (and leaves `name` alone) when one is missing
inline bool piece_name( std::string &name, const std::string &step, int rank )
{
    const auto star = name.find( '*' );
    if ( star == std::string::npos )
        return false;
    name.replace( star, 1, step );
    const auto pct = name.find( '%' );
    if ( pct == std::string::npos )
        return false;
    name.replace( pct, 1, "_" + std::to_string( rank ) );
    return true;
}

// index file naming every rank's piece (vtk_writer.h:153-199)
inline void writeParticlesParallelFile( int nranks, const std::string &step, std::string pattern )
{
    pattern.replace( pattern.find( '*' ), 1, step );
    std::string index = pattern;
    index.erase( index.find( '%' ), 1 );
    index.replace( index.find( ".vtu" ), 4, ".pvtu" );
    std::string text = "<?xml version=\"1.0\"?>\n"
                       "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" "
                       "header_type=\"UInt32\">\n"
                       "<PUnstructuredGrid>\n"
                       "\t<PPointData>\n"
                       "\t\t<PDataArray type=\"Float64\" Name=\"Velocity\"/>\n"
                       "\t\t<PDataArray type=\"Int32\" Name=\"Id\"/>\n"
                       "\t\t<PDataArray type=\"Int32\" Name=\"Type\"/>\n"
                       "\t</PPointData>\n"
                       "\t<PCellData>\n"
                       "\t</PCellData>\n"
                       "\t<PPoints>\n"
                       "\t\t<PDataArray type=\"Float64\" Name=\"Points\" NumberOfComponents=\"3\"/>\n"
                       "\t</PPoints>\n";
    for ( int r = 0; r < nranks; r++ )
    {
        std::string piece = pattern;
        piece.replace( piece.find( '%' ), 1, "_" + std::to_string( r ) );
        text += "\t<Piece Source=\"" + piece + "\"/>\n";
    }
    text += "</PUnstructuredGrid>\n</VTKFile>\n";
    if ( FILE *fp = std::fopen( index.c_str(), "w" ) )
    {
        std::fwrite( text.data(), 1, text.size(), fp );
        std::fclose( fp );
    }
}

// host copy of one rank's owned atoms at one step
struct Snapshot
{
    int n = 0;
    std::vector<double> x, v; // [n][3]
    std::vector<int> id, type;
    std::string file;
};

namespace detail
{
// appends printf("%g ", value)
inline void put_g( std::string &out, double value )
{
    char buf[40];
    const auto r = std::to_chars( buf, buf + sizeof buf, value, std::chars_format::general, 6 );
    out.append( buf, r.ptr );
    out.push_back( ' ' );
}
inline void put_d( std::string &out, int value )
{
    char buf[16];
    const auto r = std::to_chars( buf, buf + sizeof buf, value );
    out.append( buf, r.ptr );
    out.push_back( ' ' );
}

struct ArrayTag
{
    const char *type, *name, *components;
};
inline std::string open_array( const ArrayTag &t )
{
    return std::string( "\t\t<DataArray type=\"" ) + t.type + "\" Name=\"" + t.name +
           "\" NumberOfComponents=\"" + t.components + "\" format=\"ascii\">\n";
}
inline const char *close_array() { return "\n\t\t</DataArray>\n"; }

// one array body, formatted by `workers` threads over contiguous row ranges
template <class F>
void format_rows( int n, int workers, std::vector<std::string> &parts, F &&row )
{
    workers = std::max( 1, std::min( workers, n / 4096 + 1 ) );
    parts.assign( workers, std::string() );
    auto work = [&]( int w )
    {
        const int lo = (int)( (long long)n * w / workers ), hi = (int)( (long long)n * ( w + 1 ) / workers );
        std::string &s = parts[w];
        s.reserve( (size_t)( hi - lo ) * 36 );
        for ( int i = lo; i < hi; i++ )
            row( s, i );
    };
    std::vector<std::thread> pool;
    for ( int w = 1; w < workers; w++ )
        pool.emplace_back( work, w );
    work( 0 );
    for ( auto &t : pool )
        t.join();
}
} // namespace detail

// formats and writes one piece file; runs on the writer thread
inline bool write_piece( const Snapshot &s, int workers )
{
    FILE *fp = std::fopen( s.file.c_str(), "w" );
    if ( !fp )
        return false;
    auto put = [&]( const std::string &t ) { std::fwrite( t.data(), 1, t.size(), fp ); };
    std::vector<std::string> parts;
    auto body = [&]( auto &&row )
    {
        detail::format_rows( s.n, workers, parts, row );
        for ( const auto &p : parts )
            put( p );
    };
    auto triple = []( const std::vector<double> &a )
    {
        return [&a]( std::string &o, int i )
        {
            detail::put_g( o, a[3 * (size_t)i] );
            detail::put_g( o, a[3 * (size_t)i + 1] );
            detail::put_g( o, a[3 * (size_t)i + 2] );
        };
    };
    auto single = []( const std::vector<int> &a )
    { return [&a]( std::string &o, int i ) { detail::put_d( o, a[i] ); }; };

    put( "<?xml version=\"1.0\"?>\n"
         "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" "
         "header_type=\"UInt32\">\n"
         "<UnstructuredGrid>\n" );
    put( "<Piece NumberOfPoints=\"" + std::to_string( s.n ) + "\" NumberOfCells=\"0\">\n" );
    put( "\t<PointData>\n" );
    put( detail::open_array( { "Float64", "Velocity", "3" } ) );
    body( triple( s.v ) );
    put( detail::close_array() );
    put( detail::open_array( { "Int32", "Id", "1" } ) );
    body( single( s.id ) );
    put( detail::close_array() );
    put( detail::open_array( { "Int32", "Type", "1" } ) );
    body( single( s.type ) );
    put( detail::close_array() );
    put( "\t</PointData>\n\t<CellData>\n\t</CellData>\n\t<Points>\n" );
    put( detail::open_array( { "Float64", "Points", "3" } ) );
    body( triple( s.x ) );
    put( detail::close_array() );
    put( "\t</Points>\n\t<Cells>\n" );
    for ( const detail::ArrayTag &t : { detail::ArrayTag{ "Int32", "connectivity", "1" },
                                        detail::ArrayTag{ "Int32", "offsets", "1" },
                                        detail::ArrayTag{ "UInt8", "types", "1" } } )
    {
        put( detail::open_array( t ) );
        put( detail::close_array() );
    }
    put( "\t</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n" );
    return std::fclose( fp ) == 0;
}

// Background writer: one pending snapshot at most; submit() blocks only while the
// previous snapshot has not been picked up yet.
class AsyncWriter
{
    std::mutex m;
    std::condition_variable cv;
    std::unique_ptr<Snapshot> pending;
    bool busy = false, stop = false;
    int failed = 0, written = 0;
    int workers;
    std::thread thread;

    void loop()
    {
        std::unique_lock<std::mutex> lock( m );
        for ( ;; )
        {
            cv.wait( lock, [&] { return pending || stop; } );
            if ( !pending )
                return;
            std::unique_ptr<Snapshot> job = std::move( pending );
            busy = true;
            cv.notify_all();
            lock.unlock();
            const bool ok = write_piece( *job, workers );
            lock.lock();
            busy = false;
            written++;
            failed += ok ? 0 : 1;
            cv.notify_all();
        }
    }

  public:
    explicit AsyncWriter( int workers_ = 0 )
        : workers( workers_ > 0 ? workers_
                                : (int)std::min( 16u, std::max( 1u, std::thread::hardware_concurrency() ) ) )
        , thread( [this] { loop(); } )
    {
    }
    ~AsyncWriter()
    {
        {
            std::lock_guard<std::mutex> lock( m );
            stop = true;
        }
        cv.notify_all();
        thread.join();
    }
    void submit( std::unique_ptr<Snapshot> s )
    {
        std::unique_lock<std::mutex> lock( m );
        cv.wait( lock, [&] { return !pending; } );
        pending = std::move( s );
        cv.notify_all();
    }
    // waits until everything submitted is on disk; returns the number of failed files
    int drain()
    {
        std::unique_lock<std::mutex> lock( m );
        cv.wait( lock, [&] { return !pending && !busy; } );
        return failed;
    }
    int files_written()
    {
        std::lock_guard<std::mutex> lock( m );
        return written;
    }
};

// writeParticles (vtk_writer.h:205-294): snapshot now, write in the background.  `system`
// needs deep_copy_to_host() and the host mirrors x, v, id, type.
template <class t_System, class t_err>
void writeParticles( AsyncWriter &writer, int rank, int nranks, int time_step, t_System *system,
                     std::string filename, t_err &err )
{
    const std::string step = step_field( time_step );
    if ( rank == 1 )
        writeParticlesParallelFile( nranks, step, filename );
    if ( filename.find( '*' ) == std::string::npos )
        log_err( err, "VTK output file does not contain required '*'" );
    if ( filename.find( '%' ) == std::string::npos )
        log_err( err, "VTK output file does not contain required '%'" );
    if ( !piece_name( filename, step, rank ) )
        return; // non-printing rank with a bad pattern: nothing sensible to write
    system->deep_copy_to_host();
    auto snap = std::make_unique<Snapshot>();
    snap->n = system->N_local;
    const size_t n = (size_t)snap->n;
    snap->x.assign( system->x.begin(), system->x.begin() + 3 * n );
    snap->v.assign( system->v.begin(), system->v.begin() + 3 * n );
    snap->id.assign( system->id.begin(), system->id.begin() + n );
    snap->type.assign( system->type.begin(), system->type.begin() + n );
    snap->file = filename;
    writer.submit( std::move( snap ) );
}

} // namespace VTKWriter

#endif
