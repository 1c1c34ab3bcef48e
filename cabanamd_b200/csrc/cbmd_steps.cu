// cbmd_steps.cu — a run of PLAIN MD steps (no rebuild, no thermo output) in one call.
//
// The reference's step loop (cabanamd_impl.h:285-399) makes six module calls per MD step; between two
// rebuild / thermo steps they are the same launches with the same arguments every time.  For a
// system as small as input/in.lj (32 000 atoms) those launches take the device about 15 us while
// issuing them one by one takes the host 25-40 us (scripts/small_system_breakdown.py), so the
// step is launch-bound.  cbmd_md_steps runs
//     initial_integrate, update_halo, zero f, force (+ update_force for half lists), final_integrate
// nsteps times: the first step through the ordinary entry points, the second one captured into a
// CUDA graph (stream capture of the same entry points: identical kernels, arguments and order),
// the rest as graph launches.  Anything that cannot be captured (several ranks: NCCL / hub
// transport and the multi-stream overlap; NVTX ranges) runs every step through the entry points.
#include "cbmd_internal.cuh"

static int plain_step( cbmd_ctx *ctx, int half )
{
    int rc = cbmd_integrate_initial( ctx );
    if ( !rc )
        rc = cbmd_update_halo( ctx );
    if ( !rc )
        rc = cbmd_zero_force( ctx );
    if ( !rc )
        rc = cbmd_force_lj( ctx, half );
    if ( !rc && half )
        rc = cbmd_update_force( ctx );
    if ( !rc )
        rc = cbmd_integrate_final( ctx );
    return rc;
}

void cbmd_graph_release( cbmd_ctx *ctx )
{
    if ( ctx->step_graph_exec )
        cudaGraphExecDestroy( ctx->step_graph_exec );
    ctx->step_graph_exec = nullptr;
}

// capture one plain step on the context stream; returns the graph or nullptr (the caller then
// falls back to the entry points; a failed capture leaves no work behind)
static cudaGraph_t capture_step( cbmd_ctx *ctx, int half, int64_t *launches )
{
    const bool timing = ctx->timing;
    ctx->timing = false; // event pairs inside a graph cannot be read back; see cbmd_md_steps
    const int64_t before = ctx->launches;
    cudaGraph_t graph = nullptr;
    if ( cudaStreamBeginCapture( ctx->stream, cudaStreamCaptureModeThreadLocal ) != cudaSuccess )
    {
        (void)cudaGetLastError();
        ctx->timing = timing;
        return nullptr;
    }
    const int rc = plain_step( ctx, half );
    const cudaError_t e = cudaStreamEndCapture( ctx->stream, &graph );
    ctx->timing = timing;
    *launches = ctx->launches - before;
    ctx->launches = before; // counted when the graph is launched
    if ( rc != 0 || e != cudaSuccess || graph == nullptr )
    {
        (void)cudaGetLastError();
        if ( graph )
            cudaGraphDestroy( graph );
        return nullptr;
    }
    return graph;
}

extern "C" int cbmd_md_steps( cbmd_ctx *ctx, int nsteps, int half )
{
    CBMD_API_BEGIN_NOJOIN
    CBMD_REQUIRE( nsteps >= 0, "negative step count" );
    const bool graphable = ctx->graph_steps && ctx->nranks == 1 && !ctx->nvtx && nsteps >= 4;
    int done = 0;
    if ( graphable )
    {
        // step 1: the ordinary way (it also settles every lazy allocation the step needs)
        if ( plain_step( ctx, half ) != 0 )
            return 1;
        done = 1;
        // the state a captured step starts from must be the state it ends in, or replaying it would
        // not be the same as calling the entry points again: the list and the ghost plan are current,
        // a final_integrate is pending (fused into the next initial_integrate), nothing else is lazy
        const bool steady = ctx->final_pending && !ctx->halo_pending && !ctx->energy_hint && ctx->flat_halo_ok;
        int64_t per_step = 0;
        cudaGraph_t graph = steady ? capture_step( ctx, half, &per_step ) : nullptr;
        if ( graph )
        {
            bool ok = true;
            if ( ctx->step_graph_exec )
            {
                cudaGraphExecUpdateResultInfo info;
                if ( cudaGraphExecUpdate( ctx->step_graph_exec, graph, &info ) != cudaSuccess )
                {
                    (void)cudaGetLastError();
                    cbmd_graph_release( ctx );
                }
            }
            if ( !ctx->step_graph_exec &&
                 cudaGraphInstantiate( &ctx->step_graph_exec, graph, 0 ) != cudaSuccess )
            {
                (void)cudaGetLastError();
                ctx->step_graph_exec = nullptr;
                ok = false;
            }
            cudaGraphDestroy( graph );
            // the capture has advanced the host-side bookkeeping (epochs, lazy flags) by one step without
            // running it: the graph launches below are that step and the ones after it
            for ( ; ok && done < nsteps; done++ )
            {
                CBMD_CUDA( cudaGraphLaunch( ctx->step_graph_exec, ctx->stream ) );
                ctx->launches += per_step;
                ctx->graph_launches++;
                if ( ctx->timing ) // unsampled calls of the per-step buckets (cbmd_timing_get scales)
                    for ( int b : { CBMD_T_FORCE, CBMD_T_FORCE_KERNEL, CBMD_T_COMM, CBMD_T_INTEGRATE } )
                        ctx->bucket[b].calls++;
            }
            if ( !ok )
            {
                // the captured step never ran, yet its bookkeeping did: positions and velocities are one
                // step behind only in the sense that the step still has to be executed — do it now
                if ( plain_step( ctx, half ) != 0 )
                    return 1;
                done++;
            }
            // no further bookkeeping: every cache keyed by the epochs compares for equality with the
            // CURRENT epoch, and the captured step has left them exactly as a step through the entry
            // points does (mirror parts current if the step keeps them current, energy cache invalid,
            // sum(m v^2) cache invalid); replaying the step does not change which of them hold
        }
    }
    for ( ; done < nsteps; done++ )
        if ( plain_step( ctx, half ) != 0 )
            return 1;
    CBMD_API_END
}
