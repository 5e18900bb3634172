#include "teuchos_mock.hpp"
