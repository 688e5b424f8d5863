// engine_g2.cu — G2 (coordinates in Fq2) instantiation of the engine.
#include "engine_impl.cuh"

namespace b200 {
namespace eng {

B200_INSTANTIATE_GROUP(Fq2)

}  // namespace eng
}  // namespace b200
