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
// falls back to the entry points; a failed capture leaves no work behind).  The regions of the
// captured step are counted as unsampled plain-step regions (timing_mode 2: event pairs inside a
// graph could not be read back); calls[b] receives how many regions of bucket b one step has.
static cudaGraph_t capture_step( cbmd_ctx *ctx, int half, int64_t *launches, int64_t calls[CBMD_T_NBUCKETS] )
{
    const int64_t before = ctx->launches;
    for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
        calls[b] = ctx->bucket[b].calls[1];
    cudaGraph_t graph = nullptr;
    if ( cudaStreamBeginCapture( ctx->stream, cudaStreamCaptureModeThreadLocal ) != cudaSuccess )
    {
        (void)cudaGetLastError();
        return nullptr;
    }
    ctx->timing_mode = 2;
    const int rc = plain_step( ctx, half );
    ctx->timing_mode = 0;
    const cudaError_t e = cudaStreamEndCapture( ctx->stream, &graph );
    // the capture ran the host side of one step without running the step: its launches and regions are
    // counted when (and as often as) the graph is launched
    *launches = ctx->launches - before;
    ctx->launches = before;
    for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
    {
        const int64_t d = ctx->bucket[b].calls[1] - calls[b];
        ctx->bucket[b].calls[1] -= d;
        calls[b] = d;
    }
    if ( rc != 0 || e != cudaSuccess || graph == nullptr )
    {
        (void)cudaGetLastError();
        if ( graph )
            cudaGraphDestroy( graph );
        return nullptr;
    }
    return graph;
}

// one plain step through the entry points, its regions sampled (mode 1) or only counted (mode 2)
static int eager_step( cbmd_ctx *ctx, int half, int mode )
{
    ctx->timing_mode = mode;
    const int rc = plain_step( ctx, half );
    ctx->timing_mode = 0;
    return rc;
}

extern "C" int cbmd_md_steps( cbmd_ctx *ctx, int nsteps, int half )
{
    CBMD_API_BEGIN_NOJOIN
    CBMD_REQUIRE( nsteps >= 0, "negative step count" );
    const bool graphable = ctx->graph_steps && ctx->nranks == 1 && !ctx->nvtx;
    bool have_graph = false, tried = false;
    int64_t per_step = 0, calls[CBMD_T_NBUCKETS] = { 0 };
    for ( int done = 0; done < nsteps; done++ )
    {
        // timers: the first 16 plain steps and one in timing_stride afterwards are timed, through the
        // entry points; the others only count their regions (cbmd_timing_get scales)
        const bool sample = ctx->timing && ( ctx->plain_seen < 16 || ctx->plain_seen % ctx->timing_stride == 0 );
        ctx->plain_seen++;
        if ( sample || !graphable || done == 0 )
        {
            // (the first step of a stretch always runs the ordinary way: it settles every lazy
            // allocation and re-split a rebuild step may have left behind)
            if ( eager_step( ctx, half, sample ? 1 : 2 ) != 0 )
                return 1;
            continue;
        }
        if ( !tried && nsteps - done >= 2 )
        {
            tried = true;
            // the state a captured step starts from must be the state it ends in, or replaying it would
            // not be the same as calling the entry points again: the list and the ghost plan are current,
            // a final_integrate is pending (fused into the next initial_integrate), nothing else is lazy
            const bool steady = ctx->final_pending && !ctx->halo_pending && !ctx->energy_hint && ctx->flat_halo_ok;
            cudaGraph_t graph = steady ? capture_step( ctx, half, &per_step, calls ) : nullptr;
            if ( graph )
            {
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
                }
                cudaGraphDestroy( graph );
                have_graph = ctx->step_graph_exec != nullptr;
                // The capture has advanced the host-side bookkeeping (epochs, lazy flags) by one step
                // without running it.  Every cache keyed by the epochs compares for equality with the
                // CURRENT epoch, so the state is that of a step through the entry points whether the
                // step is now run from the graph or (no graph) once more through the entry points.
            }
        }
        if ( have_graph )
        {
            CBMD_CUDA( cudaGraphLaunch( ctx->step_graph_exec, ctx->stream ) );
            ctx->launches += per_step;
            ctx->graph_launches++;
            if ( ctx->timing )
                for ( int b = 0; b < CBMD_T_NBUCKETS; b++ )
                    ctx->bucket[b].calls[1] += calls[b];
        }
        else if ( eager_step( ctx, half, 2 ) != 0 )
            return 1;
    }
    CBMD_API_END
}
