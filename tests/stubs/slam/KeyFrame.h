// TEST STUB: ORB_SLAM::KeyFrame is only named (pointer parameters) by the facade header.
#pragma once
namespace ORB_SLAM { class KeyFrame; }
