#pragma once
#include <utility>
namespace absl { using std::index_sequence; using std::make_index_sequence; using std::index_sequence_for; using std::integer_sequence;
using std::in_place_t; using std::in_place; using std::exchange; using std::move; using std::forward; }
