// relations.h -- what engine.cu needs to know about relations.cu
#pragma once

struct colibri_b200_rindex;

namespace colibri {
// colibri_b200_rindex_cooc_of keeps the counters of the last pattern asked for, per index (a caller asks twice: for the number of relations, then
// for the relations); colibri_b200_rindex_free drops them
void rindex_cooc_forget(const colibri_b200_rindex* r);
}  // namespace colibri
