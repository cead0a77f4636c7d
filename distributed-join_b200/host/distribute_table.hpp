// distribute_table.hpp -- scatter a table from the root rank / gather per-rank tables on the root,
// with the reference's signatures (src/distribute_table.hpp:38-55).  Used by the tests that compare
// a distributed result with a single-GPU one (test/compare_against_single_gpu.cu:120-128,163).
#pragma once

#include <memory>

#include "communicator.hpp"
#include "cudf_shim.hpp"

// Collective.  `global_table` is only significant on rank 0; rank r receives the contiguous slice
// of size / N rows (+1 for the first size % N ranks) that src/distribute_table.cpp:39-49 assigns it.
std::unique_ptr<cudf::table> distribute_table(cudf::table_view global_table, Communicator* communicator);

// Collective.  Concatenates every rank's table on rank 0 in rank order; nullptr elsewhere.
std::unique_ptr<cudf::table> collect_tables(cudf::table_view table, Communicator* communicator);
