// decode_ring.cuh -- what the C ABI layer needs to know about the TMA ring decode kernel's scratch.
#pragma once

namespace b200 {
// ticket counter of k_decode_filter_ring (one int, zeroed before every launch); the block is padded to
// its own 128 B line so the atomics do not share a line with the slab cursors
static constexpr int kTicketInts = 32;
}  // namespace b200
