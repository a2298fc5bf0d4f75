// engine_d.cu -- fp64 instantiation of the engine (the parity-grade path: <= 1e-13 vs the reference Dslash).
#include "engine_impl.cuh"
namespace b200 {
EngineBase* make_engine_double(const Config& c) { return new Engine<double>(c); }
}  // namespace b200
