#pragma once
namespace eva { namespace msg { struct CKKSParameters {}; struct CKKSSignature {}; } }
