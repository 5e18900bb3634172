#pragma once
#include "teuchos_mock.hpp"   // the real header drags Teuchos in; FEDDLib's SmallMatrix.hpp relies on that
namespace KokkosClassic { struct DefaultNode { struct DefaultNodeType { }; }; }
