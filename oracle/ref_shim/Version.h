#pragma once
#define GITHASH "reference tree at /root/reference (no git metadata)"
