// cabanamd.h — the application object: module wiring (init) and the MD step loop (run)
// in the call order of the reference (src/cabanamd.h:65-104, src/cabanamd_impl.h:70-432,
// 644-676), with its output format.
#ifndef CBMD_HOST_CABANAMD_H
#define CBMD_HOST_CABANAMD_H

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <memory>
#include <string>

#include "binary_dump.h"
#include "binning.h"
#include "comm.h"
#include "force.h"
#include "inputCL.h"
#include "inputFile.h"
#include "integrator_nve.h"
#include "neighbor.h"
#include "output.h"
#include "property.h"
#include "read_data.h"
#include "system.h"
#include "vtk_writer.h"

class CabanaMD
{
  public:
    int nsteps = 0;
    virtual ~CabanaMD() {}
    virtual void init( InputCL cl ) = 0;
    virtual void run() = 0;
    virtual void dump_binary( int ) = 0;
    virtual void check_correctness( int ) = 0;
};

template <class t_System, class t_Neighbor>
class CbnMD : public CabanaMD
{
  public:
    bool _print_lammps = false;

    t_System *system = nullptr;
    t_Neighbor *neighbor = nullptr;
    Force<t_System, t_Neighbor> *force = nullptr;
    Integrator<t_System> *integrator = nullptr;
    Comm<t_System> *comm = nullptr;
    InputFile<t_System> *input = nullptr;
    Binning<t_System> *binning = nullptr;
    std::unique_ptr<VTKWriter::AsyncWriter> vtk; // created on the first particle dump

    ~CbnMD() override
    {
        delete force;
        delete neighbor;
        delete binning;
        delete integrator;
        delete comm;
        delete input;
        delete system;
    }

    void init( InputCL commandline ) override
    {
        system = new t_System;
        // the deck is parsed before the device context is created so that input errors
        // are reported even on a machine without a GPU (the reference creates the
        // System first; the order is not observable otherwise)
        input = new InputFile<t_System>( commandline, system );
        input->read_file();
        system->init();
        nsteps = input->nsteps;
        std::ofstream out( input->output_file, std::ofstream::app );
        std::ofstream err( input->error_file, std::ofstream::app );
        log( out, "Read input file." );
        log( out, "  CUDA (sm_100a) execution space: ", cbmd_version(), ", device ", World::get().device );

        if ( input->force_type == FORCE_NNP )
            log_err( err, "NNP requested, but not compiled!" );

        const auto neigh_cutoff = input->force_cutoff + input->neighbor_skin;
        const bool half_neigh = input->force_iteration_type == FORCE_ITER_NEIGH_HALF;

        system->sync_parameters();
        comm = new Comm<t_System>( system, neigh_cutoff );
        integrator = new Integrator<t_System>( system );
        binning = new Binning<t_System>( system );
        neighbor = new t_Neighbor( neigh_cutoff, half_neigh, input->max_neigh_guess );
        if ( input->force_type == FORCE_LJ )
            force = new ForceLJ<t_System, t_Neighbor>( system );
        else
            log_err( err, "Invalid ForceType" );
        if ( !force ) // non-printing ranks do not throw in log_err
            throw std::runtime_error( "Invalid ForceType" );
        force->init_coeff( input->force_coeff_lines );

        log( out, "Using: SystemVectorLength: ", 1, " ", system->name() );
        log( out, "Using: ", force->name(), " ", neighbor->name(), " ", comm->name(), " ", binning->name(),
             " ", integrator->name() );

        // atoms: LAMMPS data file or fcc/sc lattice (cabanamd_impl.h:185-194)
        if ( system->N == 0 && input->read_data_flag )
        {
            const int ntypes_before = system->ntypes;
            read_lammps_data_file( input, system, comm );
            if ( system->ntypes != ntypes_before )
            {
                // the type count came from the file header: size the pair tables for it
                // (the reference sizes them once, before any atoms exist)
                delete force;
                force = new ForceLJ<t_System, t_Neighbor>( system );
                force->init_coeff( input->force_coeff_lines );
            }
        }
        else if ( system->N == 0 )
            input->create_lattice( comm );
        log( out, "Created atoms." );

        comm->create_domain_decomposition();
        comm->exchange();
        binning->create_binning( neigh_cutoff, neigh_cutoff, neigh_cutoff, 1, true, false, true );
        comm->exchange_halo();
        neighbor->create( system );

        cbmd_check( cbmd_zero_force( system->ctx ), "cbmd_zero_force" );
        if ( input->thermo_rate > 0 )
            force->energy_follows();
        force->compute( system, neighbor );
        if ( half_neigh )
            comm->update_force();

        if ( input->thermo_rate > 0 )
        {
            Temperature<t_System> temp( comm );
            PotE<t_System, t_Neighbor> pote( comm );
            KinE<t_System> kine( comm );
            const auto T = temp.compute( system );
            const auto PE = pote.compute( system, force, neighbor ) / system->N;
            const auto KE = kine.compute( system ) / system->N;
            print_summary( out, 0, T, PE, KE, 0.0, 0.0 );
        }

        if ( input->dumpbinaryflag )
            dump_binary( 0 );
        if ( input->correctnessflag )
            check_correctness( 0 );
    }

    void run() override
    {
        std::ofstream out( input->output_file, std::ofstream::app );
        std::ofstream err( input->error_file, std::ofstream::app );

        const auto neigh_cutoff = input->force_cutoff + input->neighbor_skin;
        const bool half_neigh = input->force_iteration_type == FORCE_ITER_NEIGH_HALF;
        const int thermo_rate = input->thermo_rate;

        Temperature<t_System> temp( comm );
        PotE<t_System, t_Neighbor> pote( comm );
        KinE<t_System> kine( comm );

        // device-time buckets (CUDA events inside the library) stand in for the
        // reference's host timers around fenced calls
        cbmd_check( cbmd_timing_enable( system->ctx, 1 ), "cbmd_timing_enable" );
        cbmd_check( cbmd_timing_reset( system->ctx ), "cbmd_timing_reset" );
        cbmd_check( cbmd_sync( system->ctx ), "cbmd_sync" );
        using clock = std::chrono::steady_clock;
        const auto t0 = clock::now();
        auto seconds = [&] { return std::chrono::duration<double>( clock::now() - t0 ).count(); };
        double last_time = 0;

        // steps with nothing but the six module calls (no rebuild, no thermo line, no dump) are handed to
        // the library in stretches: cbmd_md_steps makes the same calls in the same order and, on one rank,
        // replays them from a CUDA graph (a step of in.lj as shipped is launch-bound otherwise).
        // CBMD_BATCH_STEPS=0 keeps every step in this loop.
        const char *be = std::getenv( "CBMD_BATCH_STEPS" );
        const bool batch_plain = !( be && std::atoi( be ) == 0 );
        auto due = []( bool on, int rate, int s ) { return on && rate > 0 && s % rate == 0; };
        auto special = [&]( int s ) {
            return s % input->comm_exchange_rate == 0 || due( true, thermo_rate, s ) ||
                   due( true, input->vtk_rate, s ) || due( input->dumpbinaryflag, input->dumpbinary_rate, s ) ||
                   due( input->correctnessflag, input->correctness_rate, s );
        };

        for ( int step = 1; step <= nsteps; step++ )
        {
            if ( batch_plain && !special( step ) )
            {
                int k = 1;
                while ( step + k <= nsteps && !special( step + k ) )
                    k++;
                cbmd_check( cbmd_md_steps( system->ctx, k, half_neigh ? 1 : 0 ), "cbmd_md_steps" );
                step += k - 1;
                continue;
            }
            integrator->initial_integrate( system );

            if ( step % input->comm_exchange_rate == 0 && step > 0 )
            {
                comm->exchange();
                binning->create_binning( neigh_cutoff, neigh_cutoff, neigh_cutoff, 1, true, false, true );
                comm->exchange_halo();
                neighbor->create( system );
            }
            else
                comm->update_halo();

            cbmd_check( cbmd_zero_force( system->ctx ), "cbmd_zero_force" );
            const bool thermo_step = thermo_rate > 0 && step % thermo_rate == 0;
            if ( thermo_step )
                force->energy_follows();
            force->compute( system, neighbor );
            if ( half_neigh )
                comm->update_force();

            integrator->final_integrate( system );

            if ( thermo_step )
            {
                const auto T = temp.compute( system );
                const auto PE = pote.compute( system, force, neighbor ) / system->N;
                const auto KE = kine.compute( system ) / system->N;
                const double time = seconds();
                const double rate = 1.0 * system->N * thermo_rate / ( time - last_time );
                print_summary( out, step, T, PE, KE, time, rate );
                last_time = time;
            }
            // vtk_rate == 0 (no `dump` line) means never; the reference takes step % 0 here
            if ( input->vtk_rate > 0 && step % input->vtk_rate == 0 )
            {
                if ( !vtk )
                    vtk = std::make_unique<VTKWriter::AsyncWriter>();
                VTKWriter::writeParticles( *vtk, comm->process_rank(), comm->num_processes(), step, system,
                                           input->vtk_file, err );
            }
            if ( input->dumpbinaryflag )
                dump_binary( step );
            if ( input->correctnessflag )
                check_correctness( step );
        }
        if ( vtk && vtk->drain() > 0 )
            log( err, "Warning: ", vtk->drain(), " VTK particle file(s) could not be written" );

        cbmd_check( cbmd_sync( system->ctx ), "cbmd_sync" );
        const double time = seconds();
        double bucket[CBMD_T_NBUCKETS];
        for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
        {
            double ms = 0;
            cbmd_check( cbmd_timing_get( system->ctx, b, &ms, nullptr ), "cbmd_timing_get" );
            bucket[b] = ms * 1e-3;
        }
        cbmd_check( cbmd_timing_enable( system->ctx, 0 ), "cbmd_timing_enable" );
        const double force_time = bucket[CBMD_T_FORCE], neigh_time = bucket[CBMD_T_NEIGH],
                     comm_time = bucket[CBMD_T_COMM], integrate_time = bucket[CBMD_T_INTEGRATE],
                     lb_time = 0.0;
        // everything not attributed above (sort, thermo, host work) is "other"
        const double other_time =
            std::max( 0.0, time - force_time - neigh_time - comm_time - integrate_time );

        if ( !_print_lammps )
        {
            const double steps_per_sec = 1.0 * nsteps / time;
            const double atom_steps_per_sec = system->N * steps_per_sec;
            const int np = comm->num_processes();
            log( out, std::fixed, std::setprecision( 2 ),
                 "\n#Procs Atoms | Time T_Force T_Neigh T_Comm T_Int T_lb ", "T_Other |\n", np, " ", system->N,
                 " | ", time, " ", force_time, " ", neigh_time, " ", comm_time, " ", integrate_time, " ",
                 lb_time, " ", other_time, " | PERFORMANCE\n", std::fixed, np, " ", system->N, " | ", 1.0, " ",
                 force_time / time, " ", neigh_time / time, " ", comm_time / time, " ",
                 integrate_time / time, " ", lb_time / time, " ", other_time / time, " | FRACTION\n\n",
                 "#Steps/s Atomsteps/s Atomsteps/(proc*s)\n", std::scientific, steps_per_sec, " ",
                 atom_steps_per_sec, " ", atom_steps_per_sec / np );
        }
        else
            log( out, "Loop time of ", time, " on ", comm->num_processes(), " procs for ", nsteps,
                 " steps with ", system->N, " atoms" );

        if ( input->write_data_flag )
        {
            system->deep_copy_to_host();
            write_data( system, comm, input->output_data_file, input->write_data_precision );
        }
    }

    // cabanamd_impl.h:434-474
    void dump_binary( int step ) override
    {
        if ( input->dumpbinary_rate <= 0 || step % input->dumpbinary_rate )
            return;
        std::ofstream err( input->error_file, std::ofstream::app );
        system->deep_copy_to_host();
        const std::string file = BinaryDump::file_name( input->dumpbinary_path, step, comm->process_rank() );
        if ( !BinaryDump::write( file, system->N_local, system->id.data(), system->type.data(),
                                 system->q.data(), system->x.data(), system->v.data(), system->f.data() ) )
        {
            log_err( err, "Cannot open dump file: ", file );
            throw std::runtime_error( "Cannot open dump file: " + file );
        }
    }

    // cabanamd_impl.h:484-642
    void check_correctness( int step ) override
    {
        if ( input->correctness_rate <= 0 || step % input->correctness_rate )
            return;
        std::ofstream err( input->error_file, std::ofstream::app );
        system->deep_copy_to_host();
        const std::string file = BinaryDump::file_name( input->reference_path, step, comm->process_rank() );
        BinaryDump::State ref;
        const char *problem = nullptr;
        switch ( BinaryDump::read( file, system->N_local, ref ) )
        {
        case BinaryDump::READ_CANNOT_OPEN:
            problem = "Cannot open input file: ";
            break;
        case BinaryDump::READ_COUNT_MISMATCH:
            problem = "Mismatch in current and reference atom counts: ";
            break;
        case BinaryDump::READ_SHORT:
            problem = "Error reading reference data: ";
            break;
        default:
            break;
        }
        if ( problem )
        {
            log_err( err, problem, file );
            throw std::runtime_error( problem + file );
        }
        BinaryDump::Deltas d = BinaryDump::compare( system->N_local, system->id.data(), system->x.data(),
                                                    system->v.data(), system->f.data(), ref );
        if ( d.unmatched_id >= 0 )
            log( err, "Unable to find current id matching reference id: ", d.unmatched_id );
        comm->reduce_float( d.sumsq, 3 );
        comm->reduce_max_float( d.maxabs, 3 );
        if ( comm->process_rank() == 0 && !BinaryDump::append_report( input->correctness_file, step, d.sumsq, d.maxabs ) )
            log( err, "Warning: cannot write correctness file ", input->correctness_file );
    }

    void print_summary( std::ofstream &out, int step, T_V_FLOAT T, T_F_FLOAT PE, T_V_FLOAT KE, double time,
                        double rate, T_INT = -1 )
    {
        if ( !_print_lammps )
        {
            if ( step == 0 )
                log( out, "\n#Timestep Temperature PotE ETot Time Atomsteps/s " );
            log( out, step, "\t", std::fixed, std::setprecision( 6 ), T, "\t", PE, "\t", PE + KE, "\t",
                 std::setprecision( 2 ), time, "\t", std::scientific, rate );
        }
        else
            log( out, "\nStep Temp E_pair TotEng CPU\n", step, " ", T, " ", PE, " ", PE + KE, " ", time );
    }
};

#endif
