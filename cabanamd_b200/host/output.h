// output.h — rank-0 logging helpers with the reference's conventions
// (src/output.h:22-54): log() writes all arguments then a newline on the printing
// rank only; log_err() does the same and then throws std::runtime_error on that rank.
#ifndef CBMD_HOST_OUTPUT_H
#define CBMD_HOST_OUTPUT_H

#include <ostream>
#include <stdexcept>
#include <utility>

// true on the rank that owns the log files (rank 0); set by Comm / main
bool print_rank();
void set_print_rank( bool is_rank0 );

template <class t_stream, class... t_args>
void log( t_stream &stream, t_args &&...args )
{
    if ( !print_rank() )
        return;
    ( stream << ... << std::forward<t_args>( args ) );
    stream << std::endl;
}

template <class t_stream, class... t_args>
void log_err( t_stream &stream, t_args &&...args )
{
    if ( !print_rank() )
        return;
    ( stream << ... << std::forward<t_args>( args ) );
    stream << std::endl;
    throw std::runtime_error( "Aborting after error from input. See error file." );
}

#endif
