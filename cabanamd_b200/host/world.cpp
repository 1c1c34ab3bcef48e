#include "world.h"

#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <thread>
#include <tuple>

#include "../../include/cbmd_c_api.h"

World &World::get()
{
    static World w;
    return w;
}

static int env_int( const char *name, int dflt )
{
    const char *v = std::getenv( name );
    return ( v && *v ) ? std::atoi( v ) : dflt;
}

void World::init_from_env()
{
    rank = env_int( "RANK", 0 );
    nranks = env_int( "WORLD_SIZE", 1 );
    device = env_int( "LOCAL_RANK", env_int( "CBMD_DEVICE", 0 ) );
    const char *f = std::getenv( "CBMD_NCCL_ID_FILE" );
    if ( f && *f )
        id_file = f;
    else
    {
        // per-user, per-launch name: torchrun's run id (when present) keeps two launches that
        // reuse a port apart; rank 0 also removes the file once every rank has joined
        const char *port = std::getenv( "MASTER_PORT" );
        const char *run = std::getenv( "TORCHELASTIC_RUN_ID" );
        id_file = std::string( "/tmp/cbmd_nccl_id_" ) + std::to_string( (long)getuid() ) + "_" +
                  ( port ? port : "0" ) + "_" + std::to_string( nranks );
        if ( run && *run )
            id_file += std::string( "_" ) + run;
    }
    if ( rank < 0 || rank >= nranks )
        throw std::runtime_error( "World: RANK outside [0, WORLD_SIZE)" );
}

// file payload: the 128-byte id followed by the publisher's wall-clock time (seconds)
static long long wall_seconds()
{
    return std::chrono::duration_cast<std::chrono::seconds>( std::chrono::system_clock::now().time_since_epoch() )
        .count();
}
static const long long ID_MAX_AGE_S = 300; // a leftover of a crashed earlier launch is ignored

void World::exchange_unique_id( unsigned char id[128] ) const
{
    if ( nranks == 1 )
    {
        std::memset( id, 0, 128 );
        return;
    }
    if ( rank == 0 )
    {
        if ( cbmd_comm_unique_id( id ) != 0 )
            throw std::runtime_error( std::string( "cbmd_comm_unique_id: " ) + cbmd_last_error() );
        // publish atomically: exclusive temporary (never follows a planted symlink), then rename
        const std::string tmp = id_file + ".tmp." + std::to_string( (long)getpid() );
        ::unlink( tmp.c_str() );
        const int fd = ::open( tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600 );
        if ( fd < 0 )
            throw std::runtime_error( "cannot create " + tmp );
        const long long now = wall_seconds();
        const bool ok = ::write( fd, id, 128 ) == 128 && ::write( fd, &now, sizeof now ) == (ssize_t)sizeof now;
        ::close( fd );
        if ( !ok )
            throw std::runtime_error( "cannot write " + tmp );
        if ( std::rename( tmp.c_str(), id_file.c_str() ) != 0 )
            throw std::runtime_error( "cannot publish " + id_file );
        return;
    }
    for ( int tries = 0; tries < 6000; tries++ ) // up to 60 s
    {
        std::ifstream in( id_file, std::ios::binary );
        if ( in )
        {
            long long stamp = 0;
            in.read( reinterpret_cast<char *>( id ), 128 );
            const bool have_id = in.gcount() == 128;
            in.read( reinterpret_cast<char *>( &stamp ), sizeof stamp );
            if ( have_id && in.gcount() == (std::streamsize)sizeof stamp && wall_seconds() - stamp <= ID_MAX_AGE_S )
                return;
        }
        std::this_thread::sleep_for( std::chrono::milliseconds( 10 ) );
    }
    throw std::runtime_error( "timed out waiting for the NCCL unique id in " + id_file );
}

// rank 0, after cbmd_comm_init has returned (ncclCommInitRank is collective: every rank has
// read the id by then): a later launch with the same file name must not find this id
void World::retire_unique_id() const
{
    if ( nranks > 1 && rank == 0 )
        ::unlink( id_file.c_str() );
}

std::array<int, 3> dims_create( int n )
{
    std::array<int, 3> best = { n, 1, 1 };
    for ( int a = 1; a <= n; a++ )
    {
        if ( n % a )
            continue;
        for ( int b = 1; b <= n / a; b++ )
        {
            if ( ( n / a ) % b )
                continue;
            int t[3] = { a, b, n / a / b };
            // sort non-increasing
            if ( t[0] < t[1] )
                std::swap( t[0], t[1] );
            if ( t[1] < t[2] )
                std::swap( t[1], t[2] );
            if ( t[0] < t[1] )
                std::swap( t[0], t[1] );
            if ( std::make_tuple( t[0] - t[2], t[0] ) < std::make_tuple( best[0] - best[2], best[0] ) )
                best = { t[0], t[1], t[2] };
        }
    }
    return best;
}
