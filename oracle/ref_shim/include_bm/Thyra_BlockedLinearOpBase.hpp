#pragma once
namespace Thyra { template <class SC> class LinearOpBase; template <class SC> class BlockedLinearOpBase; }
