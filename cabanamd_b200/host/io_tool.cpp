// cbmd_io_tool — host-only driver for the file formats either side of the MD path
// (read_data.h, vtk_writer.h, binary_dump.h).  It links no CUDA code: the templates are
// instantiated over a plain host particle store, so the CPU test-suite can pin the
// formats without a device.  Not part of the product path.
//
//   cbmd_io_tool data IN OUT [precision [atom_style [xlo xhi ylo yhi zlo zhi [rank nranks]]]]
//       parse IN as one rank (optionally owning only the given sub-box), write OUT with
//       write_data, print `N N_local ntypes | masses`.  With rank/nranks the call plays ONE
//       rank of a multi-rank write_data: ranks > 0 leave their part files, rank 0 (run last)
//       assembles the file
//   cbmd_io_tool vtk IN PATTERN STEP RANK NRANKS [WORKERS]
//       parse IN, write the particle dump(s) through the background writer
//   cbmd_io_tool dump IN PATH STEP RANK          binary dump of the parsed state (f = -x)
//   cbmd_io_tool check IN PATH STEP RANK FILE    compare with PATH/output.*, append report
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "binary_dump.h"
#include "read_data.h"
#include "vtk_writer.h"

namespace
{
struct HostSystem
{
    T_INT N = 0, N_local = 0, N_ghost = 0;
    int ntypes = 1;
    std::string atom_style = "atomic";
    std::vector<T_V_FLOAT> mass = std::vector<T_V_FLOAT>( 1, 1.0 );
    T_X_FLOAT global_mesh_lo[3] = { 0, 0, 0 }, global_mesh_hi[3] = { 0, 0, 0 };
    T_X_FLOAT local_mesh_lo_x = 0, local_mesh_lo_y = 0, local_mesh_lo_z = 0;
    T_X_FLOAT local_mesh_hi_x = 0, local_mesh_hi_y = 0, local_mesh_hi_z = 0;
    std::vector<T_X_FLOAT> x;
    std::vector<T_V_FLOAT> v;
    std::vector<T_F_FLOAT> f;
    std::vector<T_INT> type, id;
    std::vector<T_FLOAT> q;
    bool own_box = false;
    double own_lo[3], own_hi[3];

    void resize( T_INT n )
    {
        x.resize( 3 * (size_t)n ), v.resize( 3 * (size_t)n ), f.resize( 3 * (size_t)n );
        type.resize( n ), id.resize( n ), q.resize( n );
    }
    void create_domain( std::array<double, 3> lo, std::array<double, 3> hi )
    {
        for ( int d = 0; d < 3; d++ )
        {
            global_mesh_lo[d] = lo[d];
            global_mesh_hi[d] = hi[d];
        }
        const double *l = own_box ? own_lo : lo.data(), *h = own_box ? own_hi : hi.data();
        local_mesh_lo_x = l[0], local_mesh_lo_y = l[1], local_mesh_lo_z = l[2];
        local_mesh_hi_x = h[0], local_mesh_hi_y = h[1], local_mesh_hi_z = h[2];
    }
    void deep_copy_to_host() {}
};

struct OneRank
{
    int rank = 0, size = 1;
    int process_rank() { return rank; }
    int num_processes() { return size; }
    void reduce_int( T_INT *, T_INT ) {}
};

bool load( const char *file, HostSystem &s )
{
    std::ifstream in( file );
    if ( !in )
    {
        std::fprintf( stderr, "cannot open %s\n", file );
        return false;
    }
    DataFile::parse( in, &s, std::cerr );
    return true;
}
} // namespace

int main( int argc, char *argv[] )
{
    set_print_rank( true );
    try
    {
        const std::string cmd = argc > 1 ? argv[1] : "";
        HostSystem s;
        if ( cmd == "data" && argc >= 4 )
        {
            const int precision = argc > 4 ? std::atoi( argv[4] ) : 6;
            if ( argc > 5 )
                s.atom_style = argv[5];
            if ( argc > 11 )
            {
                s.own_box = true;
                for ( int d = 0; d < 3; d++ )
                {
                    s.own_lo[d] = std::atof( argv[6 + 2 * d] );
                    s.own_hi[d] = std::atof( argv[7 + 2 * d] );
                }
            }
            if ( !load( argv[2], s ) )
                return 2;
            OneRank comm;
            if ( argc > 13 )
            {
                comm.rank = std::atoi( argv[12] );
                comm.size = std::atoi( argv[13] );
            }
            write_data( &s, &comm, argv[3], precision );
            std::printf( "%d %d %d |", s.N, s.N_local, s.ntypes );
            for ( double m : s.mass )
                std::printf( " %.17g", m );
            std::printf( "\n" );
            return 0;
        }
        if ( cmd == "vtk" && argc >= 7 )
        {
            if ( !load( argv[2], s ) )
                return 2;
            VTKWriter::AsyncWriter writer( argc > 7 ? std::atoi( argv[7] ) : 0 );
            const int step = std::atoi( argv[4] ), rank = std::atoi( argv[5] ), nranks = std::atoi( argv[6] );
            // two dumps back to back exercise the hand-over to the background thread
            VTKWriter::writeParticles( writer, rank, nranks, step, &s, argv[3], std::cerr );
            VTKWriter::writeParticles( writer, rank, nranks, step + 1, &s, argv[3], std::cerr );
            return writer.drain() == 0 && writer.files_written() == 2 ? 0 : 3;
        }
        if ( ( cmd == "dump" && argc >= 6 ) || ( cmd == "check" && argc >= 7 ) )
        {
            if ( !load( argv[2], s ) )
                return 2;
            for ( size_t i = 0; i < s.f.size(); i++ )
                s.f[i] = -s.x[i];
            const int step = std::atoi( argv[4] ), rank = std::atoi( argv[5] );
            const std::string file = BinaryDump::file_name( argv[3], step, rank );
            if ( cmd == "dump" )
                return BinaryDump::write( file, s.N_local, s.id.data(), s.type.data(), s.q.data(), s.x.data(),
                                          s.v.data(), s.f.data() )
                           ? 0
                           : 3;
            BinaryDump::State ref;
            const auto st = BinaryDump::read( file, s.N_local, ref );
            if ( st != BinaryDump::READ_OK )
            {
                std::fprintf( stderr, "read status %d\n", (int)st );
                return 10 + (int)st;
            }
            const auto d =
                BinaryDump::compare( s.N_local, s.id.data(), s.x.data(), s.v.data(), s.f.data(), ref );
            if ( d.unmatched_id >= 0 )
                std::fprintf( stderr, "unmatched id %d\n", d.unmatched_id );
            return BinaryDump::append_report( argv[6], step, d.sumsq, d.maxabs ) ? 0 : 3;
        }
        std::fprintf( stderr, "usage: cbmd_io_tool data|vtk|dump|check ... (see io_tool.cpp)\n" );
        return 1;
    }
    catch ( const std::exception &e )
    {
        std::fprintf( stderr, "cbmd_io_tool: %s\n", e.what() );
        return 1;
    }
}
