// stubs.cu -- phases not implemented yet fail loudly (never a CPU fallback).
#include "engine_impl.cuh"
namespace pgb {
#ifndef PGB_HAVE_TIPS
template <int NW> void EngineT<NW>::remove_tips(TipStats*) { throw std::runtime_error("pgb200: remove_tips not implemented"); }
template void EngineT<2>::remove_tips(TipStats*);
template void EngineT<4>::remove_tips(TipStats*);
#endif
#ifndef PGB_HAVE_EDGES
template <int NW> void EngineT<NW>::build_edges(EdgeStats*, std::string*) { throw std::runtime_error("pgb200: build_edges not implemented"); }
template <int NW> void EngineT<NW>::vertices(std::string*, uint64_t*) { throw std::runtime_error("pgb200: vertices not implemented"); }
template void EngineT<2>::build_edges(EdgeStats*, std::string*);
template void EngineT<4>::build_edges(EdgeStats*, std::string*);
template void EngineT<2>::vertices(std::string*, uint64_t*);
template void EngineT<4>::vertices(std::string*, uint64_t*);
#endif
#ifndef PGB_HAVE_PASS2
template <int NW> void EngineT<NW>::pass2(Pass2Stats*, std::string*, std::string*, std::string*) { throw std::runtime_error("pgb200: pass2 not implemented"); }
template void EngineT<2>::pass2(Pass2Stats*, std::string*, std::string*, std::string*);
template void EngineT<4>::pass2(Pass2Stats*, std::string*, std::string*, std::string*);
#endif
}
