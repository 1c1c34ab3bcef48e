// inputFile.h — LAMMPS-style input deck reader and lattice / velocity initialiser behind
// the reference's InputFile surface (src/inputFile.h:150-277, src/inputFile_impl.h).
//
// The accepted command subset, the defaults, the echo of the deck into the output file
// and the error messages follow the reference; the initial state
// (fcc/sc fill order, hashed "geom" velocity RNG, momentum zeroing, rescale to the target
// temperature measured on the device) is bit-identical so both codes start from the same
// atoms (inputFile_impl.h:536-868, RNG inputFile.h:73-148).
#ifndef CBMD_HOST_INPUTFILE_H
#define CBMD_HOST_INPUTFILE_H

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "comm.h"
#include "inputCL.h"
#include "output.h"
#include "property.h"
#include "system.h"

// LAMMPS "velocity ... loop geom" generator: Park-Miller minimal standard (Schrage
// form) seeded per atom from a Jenkins one-at-a-time hash of the seed and position bytes.
class LAMMPS_RandomVelocityGeom
{
    int seed = 0;

  public:
    double uniform()
    {
        constexpr int ia = 16807, im = 2147483647, iq = 127773, ir = 2836;
        const int k = seed / iq;
        seed = ia * ( seed - k * iq ) - ir * k;
        if ( seed < 0 )
            seed += im;
        return ( 1.0 / im ) * seed;
    }
    void reset( int ibase, const double *coord )
    {
        unsigned int hash = 0;
        auto absorb = [&hash]( const void *p, size_t n )
        {
            const char *bytes = static_cast<const char *>( p ); // plain char: sign-extends on x86
            for ( size_t i = 0; i < n; i++ )
            {
                hash += bytes[i];
                hash += ( hash << 10 );
                hash ^= ( hash >> 6 );
            }
        };
        absorb( &ibase, sizeof( int ) );
        absorb( coord, 3 * sizeof( double ) );
        hash += ( hash << 3 );
        hash ^= ( hash >> 11 );
        hash += ( hash << 15 );
        seed = hash & 0x7ffffff; // 27 bits, as in the reference
        if ( !seed )
            seed = 1;
        for ( int i = 0; i < 5; i++ )
            uniform();
    }
};

inline std::vector<std::string> split( const std::string &line )
{
    std::vector<std::string> words;
    std::string cur;
    for ( char c : line )
    {
        if ( c == ' ' || c == '\t' || c == '\r' || c == '\n' )
        {
            if ( !cur.empty() )
                words.push_back( cur );
            cur.clear();
        }
        else
            cur.push_back( c );
    }
    if ( !cur.empty() )
        words.push_back( cur );
    return words;
}

template <class t_System>
class InputFile
{
    bool timestepflag = false;

  public:
    InputCL commandline;
    t_System *system;

    int units_style = UNITS_LJ;
    int lattice_style = LATTICE_FCC;
    double lattice_constant = 0.8442, lattice_offset_x = 0.0, lattice_offset_y = 0.0,
           lattice_offset_z = 0.0;

    struct Block
    {
        double xlo, xhi, ylo, yhi, zlo, zhi;
    };
    std::vector<std::string> region_order; // definition order
    std::unordered_map<std::string, Block> regions;
    std::unordered_map<std::string, int> regions_to_type;

    T_X_FLOAT min_x = std::numeric_limits<T_X_FLOAT>::max();
    T_X_FLOAT min_y = std::numeric_limits<T_X_FLOAT>::max();
    T_X_FLOAT min_z = std::numeric_limits<T_X_FLOAT>::max();
    T_X_FLOAT max_x = std::numeric_limits<T_X_FLOAT>::min();
    T_X_FLOAT max_y = std::numeric_limits<T_X_FLOAT>::min();
    T_X_FLOAT max_z = std::numeric_limits<T_X_FLOAT>::min();

    std::string output_file, error_file;

    struct Velocity
    {
        double temp = 0.0;
        int seed = 0;
    };
    std::unordered_map<int, Velocity> type_to_temperature;

    int integrator_type = INTEGRATOR_NVE;
    int nsteps = 100;
    int binning_type = BINNING_LINKEDCELL;
    int comm_type = COMM_MPI;
    int comm_exchange_rate = 20;
    double comm_ghost_cutoff;

    int force_type = FORCE_LJ;
    int force_iteration_type;
    int force_neigh_parallel_type;
    T_F_FLOAT force_cutoff = 2.5;
    std::vector<std::vector<std::string>> force_coeff_lines;

    T_F_FLOAT neighbor_skin = 0.0;
    int neighbor_type = NEIGH_VERLET_2D;
    T_INT max_neigh_guess = 50;

    int thermo_rate = 10;
    // `--dumpbinary` / `--correctness` (binary_dump.h).  The reference parses them in
    // InputCL but never copies them here; this build does.
    int dumpbinary_rate = 0, correctness_rate = 0;
    bool dumpbinaryflag = false, correctnessflag = false;
    std::string dumpbinary_path, reference_path, correctness_file;
    std::string input_data_file, output_data_file;
    int write_data_precision = 6; // stream default, as the reference writes
    bool read_data_flag = false, write_data_flag = false, write_vtk_flag = false;
    int vtk_rate = 0; // 0 = never (the reference takes step % 0 here, Appendix B.2)
    std::string vtk_file;

    InputFile( InputCL cl, t_System *s )
        : commandline( cl )
        , system( s )
    {
        neighbor_type = cl.neighbor_type;
        force_iteration_type = cl.force_iteration_type;
        force_neigh_parallel_type = cl.force_neigh_parallel_type;
        output_file = cl.output_file;
        error_file = cl.error_file;
        if ( cl.dumpbinaryflag && cl.dumpbinary_rate > 0 && cl.dumpbinary_path )
        {
            dumpbinaryflag = true;
            dumpbinary_rate = cl.dumpbinary_rate;
            dumpbinary_path = cl.dumpbinary_path;
        }
        if ( cl.correctnessflag && cl.correctness_rate > 0 && cl.reference_path && cl.correctness_file )
        {
            correctnessflag = true;
            correctness_rate = cl.correctness_rate;
            reference_path = cl.reference_path;
            correctness_file = cl.correctness_file;
        }
        // (sic) computed from the default DENSITY as if it were a lattice constant
        // (inputFile_impl.h:82-83); only sizes the Verlet bounding grid
        comm_ghost_cutoff = std::pow( 4.0 / lattice_constant, 1.0 / 3.0 ) * 20.0;
    }

    void read_file( const char *filename = nullptr )
    {
        // first use of the streams: truncate
        std::ofstream out( output_file, std::ofstream::out );
        std::ofstream err( error_file, std::ofstream::out );
        if ( !filename )
            filename = commandline.input_file;
        if ( commandline.input_file_type != INPUT_LAMMPS || !filename )
            log_err( err, "Unknown input file type: ", filename ? filename : "(none)" );
        std::ifstream in( filename );
        if ( !in )
            log_err( err, "Cannot open input file: ", filename );
        read_lammps_file( in, out, err );
    }

    void read_lammps_file( std::ifstream &in, std::ofstream &out, std::ofstream &err )
    {
        log( out, "\n#InputFile:\n", "#=========================================================" );
        std::string line;
        while ( std::getline( in, line ) )
        {
            check_lammps_command( line, err );
            log( out, line );
        }
        log( out, "#=========================================================\n" );
    }

    void check_lammps_command( std::string line, std::ofstream &err )
    {
        const auto words = split( line );
        if ( words.empty() || words[0][0] == '#' )
            return;
        const std::string &key = words[0];
        auto arg = [&]( size_t i ) -> const std::string &
        {
            if ( i >= words.size() )
                log_err( err, "LAMMPS-Command: too few arguments: ", line );
            return words.at( i );
        };
        auto num = [&]( size_t i ) { return std::stod( arg( i ) ); };
        auto integer = [&]( size_t i ) { return std::stoi( arg( i ) ); };

        if ( key == "variable" )
            log_err( err, "LAMMPS-Command: 'variable' keyword is not supported in CabanaMD" );
        else if ( key == "units" )
        {
            if ( arg( 1 ) == "metal" )
            {
                units_style = UNITS_METAL;
                system->boltz = 8.617343e-5;
                system->mvv2e = 1.0364269e-4;
                system->dt = 0.001;
            }
            else if ( arg( 1 ) == "real" )
            {
                units_style = UNITS_REAL;
                system->boltz = 0.0019872067;
                system->mvv2e = 48.88821291 * 48.88821291;
                if ( !timestepflag )
                    system->dt = 1.0;
            }
            else if ( arg( 1 ) == "lj" )
            {
                units_style = UNITS_LJ;
                system->boltz = 1.0;
                system->mvv2e = 1.0;
                if ( !timestepflag )
                    system->dt = 0.005;
            }
            else
                log_err( err, "LAMMPS-Command: 'units' command only supports 'metal', 'real', and 'lj' "
                              "in CabanaMD" );
        }
        else if ( key == "atom_style" )
        {
            if ( arg( 1 ) == "charge" )
                system->atom_style = "charge";
            else if ( arg( 1 ) != "atomic" )
                log_err( err, "LAMMPS-Command: 'atom_style' command only supports 'atomic' and 'charge' "
                              "in CabanaMD" );
        }
        else if ( key == "lattice" )
        {
            if ( arg( 1 ) == "sc" )
            {
                lattice_style = LATTICE_SC;
                lattice_constant = num( 2 );
            }
            else if ( arg( 1 ) == "fcc" )
            {
                lattice_style = LATTICE_FCC;
                // LJ units: the number is the reduced density
                lattice_constant = units_style == UNITS_LJ ? std::pow( 4.0 / num( 2 ), 1.0 / 3.0 ) : num( 2 );
            }
            else
                log_err( err, "LAMMPS-Command: 'lattice' command only supports 'sc' and 'fcc' in CabanaMD" );
            if ( words.size() > 3 )
            {
                if ( words[3] != "origin" )
                    log_err( err, "LAMMPS-Command: 'lattice' command only supports 'origin' additional "
                                  "option in CabanaMD" );
                lattice_offset_x = num( 4 );
                lattice_offset_y = num( 5 );
                lattice_offset_z = num( 6 );
            }
        }
        else if ( key == "region" )
        {
            if ( arg( 2 ) != "block" )
                log_err( err, "LAMMPS-Command: 'region' command only supports 'block' option in CabanaMD" );
            Block b{ num( 3 ), num( 4 ), num( 5 ), num( 6 ), num( 7 ), num( 8 ) };
            if ( !regions.count( arg( 1 ) ) )
                region_order.push_back( arg( 1 ) );
            regions[arg( 1 )] = b;
            min_x = std::min( min_x, lattice_constant * b.xlo );
            min_y = std::min( min_y, lattice_constant * b.ylo );
            min_z = std::min( min_z, lattice_constant * b.zlo );
            max_x = std::max( max_x, lattice_constant * b.xhi );
            max_y = std::max( max_y, lattice_constant * b.yhi );
            max_z = std::max( max_z, lattice_constant * b.zhi );
        }
        else if ( key == "create_box" )
            system->ntypes = integer( 1 );
        else if ( key == "create_atoms" )
        {
            // create_atoms TYPE box | create_atoms TYPE region REGION-ID
            if ( arg( 2 ) == "region" )
            {
                if ( !regions.count( arg( 3 ) ) )
                    log_err( err, "LAMMPS-Command: region '", arg( 3 ), "' is not defined" );
                regions_to_type[arg( 3 )] = integer( 1 );
            }
            else if ( arg( 2 ) == "box" )
            {
                if ( region_order.empty() )
                    log_err( err, "LAMMPS-Command: 'create_atoms' needs a region" );
                regions_to_type[region_order.front()] = integer( 1 );
            }
            else
                log_err( err, "LAMMPS-Command: 'create_atoms' command only supports 'region' option in "
                              "CabanaMD" );
        }
        else if ( key == "mass" )
        {
            const int t = integer( 1 ) - 1;
            if ( t < 0 )
                log_err( err, "LAMMPS-Command: 'mass' needs a positive atom type" );
            if ( t >= (int)system->mass.size() )
                system->mass.resize( t + 1, 1.0 );
            system->mass[t] = num( 2 );
        }
        else if ( key == "read_data" )
        {
            read_data_flag = true;
            input_data_file = arg( 1 );
        }
        else if ( key == "write_data" )
        {
            write_data_flag = true;
            output_data_file = arg( 1 );
            // extension: `write_data FILE precision P` (the reference reads FILE only)
            if ( words.size() > 3 && words[2] == "precision" )
                write_data_precision = std::max( 1, std::min( 17, integer( 3 ) ) );
        }
        else if ( key == "dump" )
        {
            if ( arg( 3 ) != "vtk" )
                log_err( err, "LAMMPS-Command: 'dump' command only supports 'vtk' in CabanaMD" );
            write_vtk_flag = true;
            vtk_file = arg( 5 );
            vtk_rate = (int)num( 4 );
            if ( arg( 2 ) != "all" )
                log_err( err, "LAMMPS-Command: 'dump' command only supports dumping 'all' types in "
                              "CabanaMD" );
            if ( vtk_file.find( '*' ) == std::string::npos )
                log_err( err, "LAMMPS-Command: 'dump' requires '*' in file name, so it can be replaced "
                              "by the time step in CabanaMD" );
            if ( vtk_file.find( '%' ) == std::string::npos )
                log_err( err, "LAMMPS-Command: 'dump' requires '%' in file name, so it can be replaced "
                              "by the rank in CabanaMD" );
        }
        else if ( key == "pair_style" )
        {
            if ( arg( 1 ) == "lj/cut" )
            {
                force_type = FORCE_LJ;
                force_cutoff = num( 2 );
            }
            else if ( arg( 1 ) == "nnp" )
            {
                force_type = FORCE_NNP; // rejected in CbnMD::init: not compiled
                force_coeff_lines.assign( 1, words );
            }
            else
                log_err( err, "LAMMPS-Command: 'pair_style' command only supports 'lj/cut' and 'nnp' "
                              "style in CabanaMD" );
        }
        else if ( key == "pair_coeff" )
        {
            if ( force_type == FORCE_NNP )
                force_cutoff = num( 3 );
            else
                force_coeff_lines.push_back( words );
        }
        else if ( key == "velocity" )
        {
            if ( arg( 2 ) != "create" )
                log_err( err, "LAMMPS-Command: 'velocity' command can only be used with option 'create' "
                              "in CabanaMD" );
            const int atom_type = arg( 1 ) == "all" ? 1 : integer( 1 );
            type_to_temperature[atom_type] = { num( 3 ), integer( 4 ) };
        }
        else if ( key == "neighbor" )
            neighbor_skin = num( 1 );
        else if ( key == "neigh_modify" )
        {
            for ( size_t i = 1; i < words.size(); i += 2 )
            {
                if ( words[i] == "every" )
                    comm_exchange_rate = integer( i + 1 );
                else if ( words[i] == "one" )
                    max_neigh_guess = integer( i + 1 );
                else
                    log_err( err, "LAMMPS-Command: 'neigh_modify' only supports 'every' and 'one' in "
                                  "CabanaMD" );
            }
        }
        else if ( key == "comm_modify" )
        {
            if ( arg( 1 ) != "cutoff" )
                log_err( err, "LAMMPS-Command: 'comm_modify' command only supports single cutoff "
                              "'cutoff' in CabanaMD" );
            if ( arg( 2 ) != "*" )
                log_err( err, "LAMMPS-Command: 'comm_modify' command only supported for all atom types "
                              "'*' in CabanaMD" );
            comm_ghost_cutoff = num( 3 );
        }
        else if ( key == "fix" )
        {
            if ( arg( 3 ) != "nve" )
                log_err( err, "LAMMPS-Command: 'fix' command only supports 'nve' style in CabanaMD" );
            integrator_type = INTEGRATOR_NVE;
        }
        else if ( key == "run" )
            nsteps = integer( 1 );
        else if ( key == "thermo" )
            thermo_rate = integer( 1 );
        else if ( key == "timestep" )
        {
            system->dt = num( 1 );
            timestepflag = true;
        }
        else if ( key == "newton" )
        {
            if ( commandline.set_force_iteration )
                log( err, "Warning: Overriding LAMMPS-Command: 'newton' replaced by commandline "
                          "--force-iteration" );
            else if ( arg( 1 ) == "on" )
                force_iteration_type = FORCE_ITER_NEIGH_HALF;
            else if ( arg( 1 ) == "off" )
                force_iteration_type = FORCE_ITER_NEIGH_FULL;
            else
                log_err( err, "LAMMPS-Command: 'newton' must be followed by 'on' or 'off'" );
        }
        else if ( key == "group" )
        {
            if ( arg( 2 ) != "region" )
                log_err( err, "LAMMPS-Command: 'group' command can only be used with 'region' in "
                              "CabanaMD" );
            (void)arg( 3 );
        }
        else
            log_err( err, "Unknown input file keyword: ", line );
    }

    // regions (with atoms assigned) containing a point, in definition order
    std::vector<std::string> get_regions( T_FLOAT xt, T_FLOAT yt, T_FLOAT zt ) const
    {
        std::vector<std::string> hit;
        const double a = lattice_constant;
        for ( const auto &rid : region_order )
        {
            const Block &b = regions.at( rid );
            if ( regions_to_type.count( rid ) && xt >= a * b.xlo && yt >= a * b.ylo && zt >= a * b.zlo &&
                 xt < a * b.xhi && yt < a * b.yhi && zt < a * b.zhi )
                hit.push_back( rid );
        }
        return hit;
    }

    // inputFile_impl.h:536-868
    void create_lattice( Comm<t_System> *comm )
    {
        std::ofstream out( output_file, std::ofstream::app );
        const double a = lattice_constant;

        std::array<double, 3> global_low = { min_x, min_y, min_z };
        std::array<double, 3> global_high = { max_x, max_y, max_z };
        if ( commandline.vacuum )
            for ( double &h : global_high )
                h *= commandline.vacuum_rate;
        system->create_domain( global_low, global_high, comm_ghost_cutoff );
        system->sync_parameters();

        const double lo[3] = { system->local_mesh_lo_x, system->local_mesh_lo_y, system->local_mesh_lo_z };
        const double hi[3] = { system->local_mesh_hi_x, system->local_mesh_hi_y, system->local_mesh_hi_z };
        const double mx[3] = { max_x, max_y, max_z };

        // integer cell range overlapping this rank's box (truncating conversions, as the
        // reference does)
        T_INT start[3], end[3];
        for ( int d = 0; d < 3; d++ )
        {
            start[d] = (T_INT)( lo[d] / a - 0.5 );
            end[d] = (T_INT)std::max( std::min( mx[d] / a, hi[d] / a + 0.5 ), (double)start[d] );
            if ( start[d] == end[d] )
                end[d] -= 1;
        }
        auto owned = [&]( double xt, double yt, double zt )
        {
            return xt >= lo[0] && yt >= lo[1] && zt >= lo[2] && xt < hi[0] && yt < hi[1] && zt < hi[2] &&
                   xt < mx[0] && yt < mx[1] && zt < mx[2];
        };

        std::vector<double> &hx = system->x;
        std::vector<int> &htype = system->type, &hid = system->id;
        hx.clear();
        htype.clear();

        if ( lattice_style == LATTICE_SC )
        {
            for ( T_INT iz = start[2]; iz <= end[2]; iz++ )
                for ( T_INT iy = start[1]; iy <= end[1]; iy++ )
                    for ( T_INT ix = start[0]; ix <= end[0]; ix++ )
                    {
                        const double xt = a * ( ix + lattice_offset_x ), yt = a * ( iy + lattice_offset_y ),
                                     zt = a * ( iz + lattice_offset_z );
                        if ( !owned( xt, yt, zt ) )
                            continue;
                        hx.insert( hx.end(), { xt, yt, zt } );
                        htype.push_back( std::rand() % system->ntypes );
                    }
        }
        else
        {
            double basis[4][3] = { { 0.0, 0.0, 0.0 }, { 0.5, 0.5, 0.0 }, { 0.5, 0.0, 0.5 }, { 0.0, 0.5, 0.5 } };
            for ( auto &b : basis )
            {
                b[0] += lattice_offset_x;
                b[1] += lattice_offset_y;
                b[2] += lattice_offset_z;
            }
            for ( T_INT iz = start[2]; iz <= end[2]; iz++ )
                for ( T_INT iy = start[1]; iy <= end[1]; iy++ )
                    for ( T_INT ix = start[0]; ix <= end[0]; ix++ )
                        for ( int k = 0; k < 4; k++ )
                        {
                            const double xt = a * ( 1.0 * ix + basis[k][0] ), yt = a * ( 1.0 * iy + basis[k][1] ),
                                         zt = a * ( 1.0 * iz + basis[k][2] );
                            if ( !owned( xt, yt, zt ) )
                                continue;
                            const auto rids = get_regions( xt, yt, zt );
                            if ( rids.empty() )
                                continue;
                            hx.insert( hx.end(), { xt, yt, zt } );
                            htype.push_back( regions_to_type.at( rids[std::rand() % rids.size()] ) - 1 );
                        }
        }
        const T_INT n = (T_INT)htype.size();
        system->N_local = n;
        system->N = n;
        system->N_ghost = 0;
        system->resize( n );
        comm->reduce_int( &system->N, 1 );
        // ids unique over all ranks: running index + exclusive prefix of the counts
        T_INT offset = n;
        comm->scan_int( &offset, 1 );
        for ( T_INT i = 0; i < n; i++ )
            hid[i] = i + 1 + offset - n;
        log( out, "Atoms: ", system->N, " ", system->N_local );

        // velocities: uniform in [-0.5,0.5)/sqrt(m), zero total momentum, rescale to T
        std::vector<double> &hv = system->v;
        double tot[4] = { 0.0, 0.0, 0.0, 0.0 }; // mass, px, py, pz
        for ( T_INT i = 0; i < n; i++ )
        {
            LAMMPS_RandomVelocityGeom rng;
            rng.reset( type_to_temperature[htype[i] + 1].seed, &hx[3 * (size_t)i] );
            const double m = system->mass.at( htype[i] );
            const double vx = rng.uniform() - 0.5, vy = rng.uniform() - 0.5, vz = rng.uniform() - 0.5;
            hv[3 * (size_t)i] = vx / std::sqrt( m );
            hv[3 * (size_t)i + 1] = vy / std::sqrt( m );
            hv[3 * (size_t)i + 2] = vz / std::sqrt( m );
            system->q[i] = 0.0;
            tot[0] += m;
            tot[1] += m * hv[3 * (size_t)i];
            tot[2] += m * hv[3 * (size_t)i + 1];
            tot[3] += m * hv[3 * (size_t)i + 2];
        }
        comm->reduce_float( &tot[1], 1 );
        comm->reduce_float( &tot[2], 1 );
        comm->reduce_float( &tot[3], 1 );
        comm->reduce_float( &tot[0], 1 );
        const double cx = tot[1] / tot[0], cy = tot[2] / tot[0], cz = tot[3] / tot[0];
        for ( T_INT i = 0; i < n; i++ )
        {
            hv[3 * (size_t)i] -= cx;
            hv[3 * (size_t)i + 1] -= cy;
            hv[3 * (size_t)i + 2] -= cz;
        }
        system->deep_copy_from_host();

        // temperature measured on the device, as the reference does
        Temperature<t_System> temp( comm );
        const T_V_FLOAT T = temp.compute( system );
        for ( T_INT i = 0; i < n; i++ )
        {
            const double s = std::sqrt( type_to_temperature[htype[i] + 1].temp / T );
            hv[3 * (size_t)i] *= s;
            hv[3 * (size_t)i + 1] *= s;
            hv[3 * (size_t)i + 2] *= s;
        }
        system->deep_copy_velocities_from_host();
    }
};

#endif
