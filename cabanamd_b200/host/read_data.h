// read_data.h — LAMMPS data files either side of the MD path: `read_data FILE` as the
// initial state instead of the lattice generator, `write_data FILE` for the final state
// (reference src/read_data.h:99-422, call sites cabanamd_impl.h:186-189,430-431).
//
// The accepted file subset is the reference's: one title line, a header with
// `N atoms`, `T atom types`, `lo hi xlo xhi|ylo yhi|zlo zhi` (the z line ends the header),
// then the sections `Atoms`, `Velocities`, `Masses`, `Pair Coeffs` in any order with Atoms
// before Velocities; `#` starts a comment; atom lines are `id type x y z` (atom_style
// atomic) or `id type q x y z` (charge).  Every rank scans the whole file and keeps the
// atoms with lo <= x < hi of its own sub-box (read_data.h:199-202).
//
// Deliberate differences, all supersets of the reference's behaviour on valid files:
//  * velocities are matched to atoms by id through a hash map, so the two sections need
//    not be in the same order (the reference walks both lists in lockstep, :254-262);
//    atoms without a velocity line start at rest instead of uninitialised memory;
//  * a `Pair Coeffs` section is skipped once (the reference calls its skipper twice,
//    :345-350, and so swallows the head of whatever section follows);
//  * write_data on several ranks gathers every rank's atoms and writes the GLOBAL box
//    (the reference lets rank 0 write its own sub-box and its own atoms under the global
//    count, :385-421); one rank produces the reference's file byte for byte;
//  * `write_data FILE precision P` (extension) raises the stream precision from the
//    reference's 6 significant digits so that a restart is exact with P = 17.
//
// The parser and the writer are host-only templates over the System type (they touch the
// host mirrors and the domain scalars only), so the CPU tests drive them through
// `cbmd_io_tool` without a device context.
#ifndef CBMD_HOST_READ_DATA_H
#define CBMD_HOST_READ_DATA_H

#include <array>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <istream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "output.h"
#include "types.h"

namespace DataFile
{

inline std::string strip_comment( const std::string &line ) { return line.substr( 0, line.find( '#' ) ); }

inline bool is_blank( const std::string &line )
{
    return line.find_first_not_of( " \r\t\n" ) == std::string::npos;
}

inline std::string trimmed( const std::string &s )
{
    const auto a = s.find_first_not_of( " \r\t\n" );
    if ( a == std::string::npos )
        return "";
    const auto b = s.find_last_not_of( " \r\t\n" );
    return s.substr( a, b - a + 1 );
}

// cursor over the numeric fields of one line (what the reference does with sscanf)
class Fields
{
    const char *p;

  public:
    explicit Fields( const std::string &line )
        : p( line.c_str() )
    {
    }
    bool integer( T_INT &out )
    {
        char *end = nullptr;
        const long v = std::strtol( p, &end, 10 );
        if ( end == p )
            return false;
        out = (T_INT)v;
        p = end;
        return true;
    }
    bool real( T_FLOAT &out )
    {
        char *end = nullptr;
        const double v = std::strtod( p, &end );
        if ( end == p )
            return false;
        out = v;
        p = end;
        return true;
    }
};

struct Header
{
    T_INT natoms = 0;
    int ntypes = 0;
    std::array<double, 3> low = { 0, 0, 0 }, high = { 0, 0, 0 };
};

// title line + header lines up to and including `zlo zhi`
template <class t_err>
Header read_header( std::istream &file, t_err &err )
{
    Header h;
    std::string line;
    if ( !std::getline( file, line ) )
        log_err( err, "Could not read from data file. Please check for a valid file and "
                      "ensure that file path is less than 32 characters." );
    bool have_z = false;
    while ( !have_z && std::getline( file, line ) )
    {
        line = strip_comment( line );
        if ( is_blank( line ) )
            continue;
        Fields f( line );
        T_INT count = 0;
        double lo = 0, hi = 0;
        // same precedence as the reference's if-chain (read_data.h:128-159)
        if ( line.find( "atoms" ) != std::string::npos )
        {
            if ( f.integer( count ) )
                h.natoms = count;
        }
        else if ( line.find( "atom types" ) != std::string::npos )
        {
            if ( f.integer( count ) )
                h.ntypes = count;
        }
        else
        {
            static const char *const axis_key[3] = { "xlo xhi", "ylo yhi", "zlo zhi" };
            for ( int d = 0; d < 3; d++ )
                if ( line.find( axis_key[d] ) != std::string::npos )
                {
                    if ( f.real( lo ) && f.real( hi ) )
                    {
                        h.low[d] = lo;
                        h.high[d] = hi;
                    }
                    have_z = d == 2;
                    break;
                }
        }
    }
    if ( !have_z )
        log_err( err, "Data file header ended before the 'zlo zhi' line" );
    return h;
}

// next section keyword ("" at end of file); unknown keywords are input errors
template <class t_err>
std::string next_section( std::istream &file, t_err &err )
{
    std::string line;
    while ( std::getline( file, line ) )
    {
        const std::string key = trimmed( strip_comment( line ) );
        if ( key.empty() )
            continue;
        if ( key != "Atoms" && key != "Velocities" && key != "Masses" && key != "Pair Coeffs" )
            log_err( err, "Unknown data file keyword: ", key );
        return key;
    }
    return "";
}

// first line of a section body: the reference skips EMPTY lines only (read_data.h:49-58)
inline bool first_body_line( std::istream &file, std::string &line )
{
    while ( std::getline( file, line ) )
        if ( !line.empty() && line != "\r" )
            return true;
    return false;
}

// Result of scanning a file for one rank: the atoms this rank owns, in file order.
struct LocalAtoms
{
    std::vector<T_INT> id, type;
    std::vector<T_FLOAT> q;
    std::vector<T_X_FLOAT> x; // [n][3]
    std::vector<T_V_FLOAT> v; // [n][3]
    std::unordered_map<T_INT, size_t> row_of_id;
    size_t size() const { return id.size(); }
};

template <class t_System>
void read_atoms( std::istream &file, t_System *s, LocalAtoms &mine )
{
    const bool charge = s->atom_style == "charge";
    const double lo[3] = { s->local_mesh_lo_x, s->local_mesh_lo_y, s->local_mesh_lo_z };
    const double hi[3] = { s->local_mesh_hi_x, s->local_mesh_hi_y, s->local_mesh_hi_z };
    std::string line;
    bool ok = first_body_line( file, line );
    for ( T_INT n = 0; n < s->N && ok; n++, ok = (bool)std::getline( file, line ) )
    {
        Fields f( line );
        T_INT id = 0, type = 0;
        T_FLOAT q = 0, p[3] = { 0, 0, 0 };
        bool parsed = f.integer( id ) && f.integer( type );
        if ( charge )
            parsed = parsed && f.real( q );
        parsed = parsed && f.real( p[0] ) && f.real( p[1] ) && f.real( p[2] );
        if ( !parsed )
            continue;
        bool inside = true;
        for ( int d = 0; d < 3; d++ )
            inside = inside && p[d] >= lo[d] && p[d] < hi[d];
        if ( !inside )
            continue;
        if ( type < 1 || ( s->ntypes > 0 && type > s->ntypes ) )
            throw std::runtime_error( "read_data: atom " + std::to_string( id ) + " has type " +
                                      std::to_string( type ) + " outside 1.." + std::to_string( s->ntypes ) +
                                      " ('atom types' of the header)" ); // on whichever rank owns it
        mine.row_of_id[id] = mine.size();
        mine.id.push_back( id );
        mine.type.push_back( type - 1 );
        mine.q.push_back( q );
        mine.x.insert( mine.x.end(), { p[0], p[1], p[2] } );
    }
    mine.v.assign( 3 * mine.size(), 0.0 );
}

template <class t_System>
void read_velocities( std::istream &file, t_System *s, LocalAtoms &mine )
{
    std::string line;
    bool ok = first_body_line( file, line );
    for ( T_INT n = 0; n < s->N && ok; n++, ok = (bool)std::getline( file, line ) )
    {
        Fields f( line );
        T_INT id = 0;
        T_FLOAT w[3];
        if ( !( f.integer( id ) && f.real( w[0] ) && f.real( w[1] ) && f.real( w[2] ) ) )
            continue;
        const auto it = mine.row_of_id.find( id );
        if ( it == mine.row_of_id.end() )
            continue; // another rank's atom
        for ( int d = 0; d < 3; d++ )
            mine.v[3 * it->second + d] = w[d];
    }
}

template <class t_System>
void read_masses( std::istream &file, t_System *s )
{
    s->mass.assign( (size_t)std::max( s->ntypes, 1 ), 1.0 );
    std::string line;
    bool ok = first_body_line( file, line );
    for ( int n = 0; n < s->ntypes && ok; n++, ok = (bool)std::getline( file, line ) )
    {
        Fields f( line );
        T_INT type = 0;
        T_FLOAT m = 0;
        if ( f.integer( type ) && f.real( m ) && type >= 1 && type <= s->ntypes )
            s->mass[type - 1] = m;
    }
}

template <class t_System>
void skip_pair_coeffs( std::istream &file, t_System *s )
{
    std::string line;
    bool ok = first_body_line( file, line );
    for ( int n = 1; n < s->ntypes && ok; n++ )
        ok = (bool)std::getline( file, line );
}

// Parses the stream into the System's host mirrors (rows [0, N_local)) and domain
// scalars.  Host work only; the caller uploads.
template <class t_System, class t_err>
void parse( std::istream &file, t_System *s, t_err &err )
{
    const Header h = read_header( file, err );
    s->N = h.natoms;
    s->ntypes = h.ntypes;
    s->create_domain( h.low, h.high ); // two-argument form, as read_data.h:162

    LocalAtoms mine;
    bool have_atoms = false;
    for ( std::string key = next_section( file, err ); !key.empty(); key = next_section( file, err ) )
    {
        if ( key == "Atoms" )
        {
            read_atoms( file, s, mine );
            have_atoms = true;
        }
        else if ( key == "Velocities" )
        {
            if ( !have_atoms )
                log_err( err, "Must read Atoms before Velocities" );
            read_velocities( file, s, mine );
        }
        else if ( key == "Masses" )
            read_masses( file, s );
        else
        {
            skip_pair_coeffs( file, s );
            log( err, "Warning: Ignoring potential parameters in data file. "
                      "CabanaMD only reads pair_coeff in the input file." );
        }
    }

    const T_INT n = (T_INT)mine.size();
    s->N_local = n;
    s->N_ghost = 0;
    s->resize( n );
    s->x = std::move( mine.x );
    s->v = std::move( mine.v );
    s->type = std::move( mine.type );
    s->id = std::move( mine.id );
    s->q = std::move( mine.q );
    s->f.assign( 3 * (size_t)n, 0.0 );
}

// Text of one rank's share of the two per-atom sections, in the reference's line format
// (read_data.h:409-420): `id type x y z` and `id vx vy vz`, stream-formatted.
template <class t_System>
void format_rows( const t_System *s, int precision, std::string &atoms, std::string &velocities )
{
    std::ostringstream a, w;
    a << std::setprecision( precision );
    w << std::setprecision( precision );
    for ( T_INT n = 0; n < s->N_local; n++ )
    {
        const size_t r = 3 * (size_t)n;
        a << s->id[n] << " " << s->type[n] + 1 << " " << s->x[r] << " " << s->x[r + 1] << " " << s->x[r + 2]
          << "\n";
        w << s->id[n] << " " << s->v[r] << " " << s->v[r + 1] << " " << s->v[r + 2] << "\n";
    }
    atoms = a.str();
    velocities = w.str();
}

// The whole file from already formatted per-atom sections.
inline void write_file( std::ostream &data, T_INT natoms, int ntypes, const double lo[3], const double hi[3],
                        int precision, const std::string &atoms, const std::string &velocities )
{
    static const char *const axis[3] = { "x", "y", "z" };
    data << std::setprecision( precision );
    data << "LAMMPS data file from CabanaMD\n\n";
    data << natoms << " atoms\n" << ntypes << " atom types\n\n";
    for ( int d = 0; d < 3; d++ )
        data << lo[d] << " " << hi[d] << " " << axis[d] << "lo " << axis[d] << "hi\n";
    data << "\nAtoms # atomic\n\n" << atoms << "\nVelocities\n\n" << velocities;
    data.flush();
}

} // namespace DataFile

// ---- reference entry points ------------------------------------------------------------

// read_lammps_data_file (read_data.h:299-383): parse, upload, verify the global count.
template <class t_Input, class t_System, class t_Comm>
void read_lammps_data_file( t_Input *input, t_System *s, t_Comm *comm )
{
    std::ofstream out( input->output_file, std::ofstream::app );
    std::ofstream err( input->error_file, std::ofstream::app );
    std::ifstream file( input->input_data_file );
    if ( !file )
        log_err( err, "Could not read from data file. Please check for a valid file and "
                      "ensure that file path is less than 32 characters." );
    DataFile::parse( file, s, err );

    s->sync_parameters(); // the Masses section may have changed the per-type masses
    s->deep_copy_from_host();

    T_INT natoms = s->N_local;
    comm->reduce_int( &natoms, 1 );
    if ( natoms != s->N )
        log_err( err, "Created incorrect # of atoms." );
    else
        log( out, "Atoms: ", s->N, " ", s->N_local );
    if ( natoms != s->N ) // non-printing ranks do not throw in log_err
        throw std::runtime_error( "Created incorrect # of atoms." );
}

// write_data (read_data.h:385-422).  `s` must hold current host mirrors of its owned atoms
// (deep_copy_to_host).  Ranks > 0 hand their rows to rank 0 through part files next to
// the target (one node, shared file system; comm->reduce_int is the barrier).
template <class t_System, class t_Comm>
void write_data( t_System *s, t_Comm *comm, const std::string &data_file, int precision = 6 )
{
    std::string atoms, velocities;
    DataFile::format_rows( s, precision, atoms, velocities );
    const int rank = comm->process_rank(), nranks = comm->num_processes();
    auto part = [&]( int r, const char *what )
    { return data_file + ".part" + std::to_string( r ) + "." + what; };
    if ( nranks > 1 )
    {
        if ( rank > 0 )
        {
            std::ofstream( part( rank, "atoms" ), std::ios::binary ) << atoms;
            std::ofstream( part( rank, "vel" ), std::ios::binary ) << velocities;
        }
        T_INT token = 1;
        comm->reduce_int( &token, 1 );
    }
    if ( rank != 0 )
        return;
    for ( int r = 1; r < nranks; r++ )
    {
        for ( const char *what : { "atoms", "vel" } )
        {
            std::ifstream in( part( r, what ), std::ios::binary );
            std::stringstream all;
            all << in.rdbuf();
            ( what[0] == 'a' ? atoms : velocities ) += all.str();
            std::remove( part( r, what ).c_str() );
        }
    }
    const double lo[3] = { s->global_mesh_lo[0], s->global_mesh_lo[1], s->global_mesh_lo[2] };
    const double hi[3] = { s->global_mesh_hi[0], s->global_mesh_hi[1], s->global_mesh_hi[2] };
    const double llo[3] = { s->local_mesh_lo_x, s->local_mesh_lo_y, s->local_mesh_lo_z };
    const double lhi[3] = { s->local_mesh_hi_x, s->local_mesh_hi_y, s->local_mesh_hi_z };
    std::ofstream data( data_file );
    // one rank: the local box IS the global box, written exactly as the reference does
    DataFile::write_file( data, s->N, s->ntypes, nranks > 1 ? lo : llo, nranks > 1 ? hi : lhi, precision, atoms,
                          velocities );
}

#endif
