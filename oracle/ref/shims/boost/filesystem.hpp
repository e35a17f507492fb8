// TEST INFRASTRUCTURE ONLY - compile-only stand-in (named by viewer headers the pinned paths never use)
#pragma once
