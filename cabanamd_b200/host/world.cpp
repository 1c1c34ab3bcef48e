#include "world.h"

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
        const char *port = std::getenv( "MASTER_PORT" );
        id_file = std::string( "/tmp/cbmd_nccl_id_" ) + ( port ? port : "0" ) + "_" +
                  std::to_string( nranks );
    }
    if ( rank < 0 || rank >= nranks )
        throw std::runtime_error( "World: RANK outside [0, WORLD_SIZE)" );
}

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
        // publish atomically: write a temporary, then rename
        const std::string tmp = id_file + ".tmp";
        {
            std::ofstream o( tmp, std::ios::binary | std::ios::trunc );
            o.write( reinterpret_cast<const char *>( id ), 128 );
            if ( !o )
                throw std::runtime_error( "cannot write " + tmp );
        }
        if ( std::rename( tmp.c_str(), id_file.c_str() ) != 0 )
            throw std::runtime_error( "cannot publish " + id_file );
        return;
    }
    for ( int tries = 0; tries < 6000; tries++ ) // up to 60 s
    {
        std::ifstream in( id_file, std::ios::binary );
        if ( in )
        {
            in.read( reinterpret_cast<char *>( id ), 128 );
            if ( in.gcount() == 128 )
                return;
        }
        std::this_thread::sleep_for( std::chrono::milliseconds( 10 ) );
    }
    throw std::runtime_error( "timed out waiting for the NCCL unique id in " + id_file );
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
