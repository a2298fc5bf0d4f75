// engine_f.cu -- fp32 instantiation of the engine (storage + arithmetic in float, reductions in double).
#include "engine_impl.cuh"
namespace b200 {
EngineBase* make_engine_float(const Config& c) { return new Engine<float>(c); }
}  // namespace b200
