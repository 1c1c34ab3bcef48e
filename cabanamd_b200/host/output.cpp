#include "output.h"

static bool g_print_rank = true;
bool print_rank() { return g_print_rank; }
void set_print_rank( bool is_rank0 ) { g_print_rank = is_rank0; }
