// comm.h — spatial-decomposition communication behind the reference's Comm surface
// (src/comm_mpi.h:125-138,355-357, src/comm_mpi_impl.h:52-441).  MPI is replaced by
// NCCL inside libcbmd_cuda; the six-phase logic runs on the device.
#ifndef CBMD_HOST_COMM_H
#define CBMD_HOST_COMM_H

#include "output.h"
#include "system.h"

template <class t_System>
class Comm
{
    t_System *system;
    T_X_FLOAT comm_depth;
    int proc_rank = 0, proc_size = 1;

  public:
    // comm_depth = force cutoff + skin: depth of the ghost shell (cabanamd_impl.h:101)
    Comm( t_System *s, T_X_FLOAT comm_depth_ )
        : system( s )
        , comm_depth( comm_depth_ )
    {
        init();
    }

    // joins the NCCL communicator (MPI_Comm_size/rank in the reference, :52-76)
    void init()
    {
        const World &w = World::get();
        proc_rank = w.rank;
        proc_size = w.nranks;
        system->init();
        unsigned char id[128];
        w.exchange_unique_id( id );
        cbmd_check( cbmd_comm_init( system->ctx, proc_size, proc_rank, proc_size > 1 ? id : nullptr ),
                    "cbmd_comm_init" );
        w.retire_unique_id();
        set_print_rank( proc_rank == 0 );
    }

    // the face neighbours follow from the rank grid handed over by
    // System::create_domain (comm_mpi_impl.h:78-119); nothing else to set up
    void create_domain_decomposition() {}

    // migrate / PBC-wrap owned atoms; returns the global number of migrated atoms (:191-278)
    T_INT exchange()
    {
        int sent = 0;
        cbmd_check( cbmd_exchange( system->ctx, &sent ), "cbmd_exchange" );
        system->refresh_counts();
        return sent;
    }
    // build the ghost shell (:280-367)
    void exchange_halo()
    {
        cbmd_check( cbmd_exchange_halo( system->ctx, comm_depth ), "cbmd_exchange_halo" );
        system->refresh_counts();
    }
    // refresh ghost positions (:369-408)
    void update_halo() { cbmd_check( cbmd_update_halo( system->ctx ), "cbmd_update_halo" ); }
    // fold ghost forces back into their owners (:410-441)
    void update_force() { cbmd_check( cbmd_update_force( system->ctx ), "cbmd_update_force" ); }

    // scalar collectives (:121-189), in place
    void scan_int( T_INT *vals, T_INT count )
    {
        cbmd_check( cbmd_scan_sum_int( system->ctx, vals, count ), "cbmd_scan_sum_int" );
    }
    void reduce_int( T_INT *vals, T_INT count )
    {
        cbmd_check( cbmd_reduce_sum_int( system->ctx, vals, count ), "cbmd_reduce_sum_int" );
    }
    void reduce_float( T_FLOAT *vals, T_INT count )
    {
        cbmd_check( cbmd_reduce_sum_double( system->ctx, vals, count ), "cbmd_reduce_sum_double" );
    }
    void reduce_max_int( T_INT *vals, T_INT count )
    {
        cbmd_check( cbmd_reduce_max_int( system->ctx, vals, count ), "cbmd_reduce_max_int" );
    }
    void reduce_max_float( T_FLOAT *vals, T_INT count )
    {
        cbmd_check( cbmd_reduce_max_double( system->ctx, vals, count ), "cbmd_reduce_max_double" );
    }
    // (sic) the reference's reduce_min_* use MPI_MAX (comm_mpi_impl.h:171-189); unused on the path
    void reduce_min_int( T_INT *vals, T_INT count ) { reduce_max_int( vals, count ); }
    void reduce_min_float( T_FLOAT *vals, T_INT count ) { reduce_max_float( vals, count ); }

    int process_rank() { return proc_rank; }
    int num_processes() { return proc_size; }
    const char *name() { return "Comm:CabanaMPI"; }
};

#endif
